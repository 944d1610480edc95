"""Developer aid: time the radix sort core in isolation (bce_gpu_dbg_radix, not part of the ABI)."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bce_b200 import Frontend  # noqa: E402

fe = Frontend(0)
f = fe.lib.bce_gpu_dbg_radix
f.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
m = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
for flags in (0, 0, 1):
    ms, bad = C.c_float(), C.c_int()
    rc = f(fe.h, m, 8, flags, C.byref(ms), C.byref(bad))
    gb = 8 * 24 * m / 1e9
    print(f"m={m} flags={flags} rc={rc} {ms.value:.3f} ms (8 passes + hist) -> {gb / (ms.value / 1e3):.0f} GB/s algorithmic, unsorted pairs={bad.value}")
