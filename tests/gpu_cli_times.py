"""Developer aid (run under gpurun): wall time of the command line tools on one synthetic file,
ours (`bce_b200/bce`, GPU front end + host coders) beside the unmodified reference
(`oracle/_ref/bce_ref`, SA-IS stand-in for libdivsufsort), archives compared byte for byte.
usage: python tests/gpu_cli_times.py <generator> <bytes> <seed> [out.json]"""
import filecmp
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bce_b200 import synth  # noqa: E402

kind, n, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out_path = sys.argv[4] if len(sys.argv) > 4 else None
ours, ref = ROOT / "bce_b200" / "bce", ROOT / "oracle" / "_ref" / "bce_ref"
res = {"generator": kind, "bytes": n, "seed": seed, "host_cores": os.cpu_count()}


def run(tag, *cmd, env=None):
    t = time.perf_counter()
    r = subprocess.run([str(c) for c in cmd], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
                       env=dict(os.environ, **(env or {})))
    res[tag + "_s"] = round(time.perf_counter() - t, 3)
    res[tag + "_rc"] = r.returncode


with tempfile.TemporaryDirectory() as d:
    d = Path(d)
    src = d / "in.bin"
    src.write_bytes(bytes(synth.generate(kind, n, seed)))
    run("ours_c_cold", ours, "-c", d / "a.bce", src)          # first CUDA context + allocations
    run("ours_c", ours, "-c", d / "a.bce", src)
    run("ours_d", ours, "-d", d / "a.out", d / "a.bce")
    res["ours_roundtrip"] = filecmp.cmp(src, d / "a.out", shallow=False)
    res["archive_bytes"] = (d / "a.bce").stat().st_size
    if ref.exists():
        run("ref_c", ref, "-c", d / "r.bce", src, env={"OMP_NUM_THREADS": "8"})
        res["archives_identical"] = filecmp.cmp(d / "a.bce", d / "r.bce", shallow=False)
        run("ref_d", ref, "-d", d / "r.out", d / "a.bce", env={"OMP_NUM_THREADS": "8"})
        res["ref_decodes_ours"] = filecmp.cmp(src, d / "r.out", shallow=False)
print(json.dumps(res))
if out_path:
    Path(out_path).write_text(json.dumps(res, indent=1) + "\n")
