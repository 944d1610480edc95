"""bce_b200 -- B200-native compression front end for BCE v0.4 (akamiru/bce).

Only the data-parallel path of `bce -c / -d / -s` lives here: suffix sort + BWT, wavelet
matrix, the CSE level loop and the inverse BWT, as sm_100a CUDA kernels behind the C ABI in
include/bce_gpu.h.  See DESIGN.md for the path, INTEGRATION.md for the reference-side binding.
"""
from .gpu import BceGpuError, Frontend, load_library  # noqa: F401

__all__ = ["Frontend", "BceGpuError", "load_library"]
