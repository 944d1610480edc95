"""CPU: the C-ABI libraries load and export every symbol include/*.h declares (no compute calls),
and the product never reaches into oracle/."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared(header):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bce_[a-z0-9_]+)\s*\(", text)))


def test_gpu_library_exports_every_declared_symbol():
    from bce_b200 import gpu
    lib = gpu.load_library()
    names = declared("bce_gpu.h")
    assert set(names) == set(gpu.ABI_SYMBOLS)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.bce_gpu_abi_version() == 1
    assert lib.bce_gpu_error_string(-5) == b"CSE frontier exceeded its device memory"


def test_host_library_exports_every_declared_symbol():
    from bce_b200 import host
    lib = host.load_library()
    names = [n for n in declared("bce_host.h")]
    assert set(names) == set(host.HOST_SYMBOLS)
    for n in names:
        assert hasattr(lib, n), n


def test_open_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bce_b200 import BceGpuError, Frontend
    with pytest.raises(BceGpuError) as e:
        Frontend(0)
    assert e.value.code == -7            # BCE_GPU_E_NODEVICE: no CPU fallback exists


def test_product_never_touches_the_oracle():
    bad = []
    for p in (ROOT / "bce_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".c") and "oracle" in p.read_text(errors="ignore"):
            bad.append(str(p))
    assert not bad, bad
    out = ctypes.CDLL(str(ROOT / "bce_b200" / "libbce_gpu.so"))
    assert not hasattr(out, "bceo_cse")
