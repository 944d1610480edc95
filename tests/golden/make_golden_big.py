"""Generates tests/golden/big_vectors.json: what the UNMODIFIED reference (oracle/_ref, built by
oracle/Makefile from /root/reference/bce.cpp) produces on BASELINE.json's configurations at
their full sizes.  Run once in the build container (about 40 minutes of CPU):

    python tests/golden/make_golden_big.py [name ...]

Per input: sha256 of the input, of the BWT and of the archive written by `bce_ref -c`
(bce.cpp:1403-1427), offset, C[8], the number of coder calls per stream and ref_tap.cpp's
order-sensitive call checksums (the streams themselves are tens of GB at 1 GB).  For the
`-s` configuration also the 288-byte config `bce_ref -s` wrote (bce.cpp:1384-1402) and the
archive of `bce_ref -c archive file cfg`.  The GPU tests (`tests/test_gpu_fullsize.py`)
regenerate the inputs from their seeds and compare against these values; nothing here
travels but the JSON."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from bce_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402

OUT = ROOT / "tests" / "golden" / "big_vectors.json"
ENV = dict(os.environ, OMP_NUM_THREADS=os.environ.get("OMP_NUM_THREADS", "6"), OMP_WAIT_POLICY="passive")
os.environ.update(OMP_NUM_THREADS=ENV["OMP_NUM_THREADS"], OMP_WAIT_POLICY="passive")   # the in-process tap

# name -> (generator, bytes, seed, run `-s` too)
INPUTS = {
    "markov2-1MB": ("markov2-text", 10**6, 1, False),          # configs[0]
    "enwik-100MB": ("enwik-shaped", 10**8, 2, False),          # configs[1]
    "enwik-1GB": ("enwik-shaped", 10**9, 3, False),            # configs[2]
    "mixed-256MB": ("mixed-binary", 268435456, 4, True),       # configs[3]
    "batch-128MB-seed100": ("enwik-shaped", 134217728, 100, False),   # configs[4], first file
    "batch-128MB-seed163": ("enwik-shaped", 134217728, 163, False),   # configs[4], last file
}


def sha(b):
    return hashlib.sha256(b).hexdigest()


def sha_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            blk = f.read(1 << 24)
            if not blk:
                break
            h.update(blk)
    return h.hexdigest()


def one(name):
    kind, n, seed, scan = INPUTS[name]
    data = synth.generate(kind, n, seed)
    v = dict(name=name, generator=kind, n=n, seed=seed, input_sha256=sha(data.tobytes()))
    with tempfile.TemporaryDirectory() as d:
        src, arc = os.path.join(d, "in"), os.path.join(d, "a.bce")
        data.tofile(src)
        # one after the other, passive OpenMP waits: two spinning 8-thread teams on 8 cores turn the
        # reference's per-round barriers (24 k rounds on the mixed input) into scheduler quanta
        t0 = time.time()
        subprocess.run([str(oracle.REF_BIN), "-c", arc, src], check=True, capture_output=True, env=ENV)
        cli = {"seconds": time.time() - t0}
        r = oracle.ref_front(data, want_bwt=True, record=False, checksum=True)
        v.update(offset=r["offset"], C=r["C"], bwt_sha256=sha(r["bwt"].tobytes()), stream_calls=r["calls"],
                 stream_checksum=[[f"{a:016x}", f"{b:016x}"] for a, b in r["checksum"]],
                 archive_bytes=os.path.getsize(arc), archive_sha256=sha_file(arc),
                 ref_cli_seconds=round(cli["seconds"], 1), ref_cli_threads=int(ENV["OMP_NUM_THREADS"]))
        if scan:
            cfg = os.path.join(d, "cfg")
            arc2 = os.path.join(d, "b.bce")
            subprocess.run([str(oracle.REF_BIN), "-s", cfg, src], check=True, capture_output=True, env=ENV)
            subprocess.run([str(oracle.REF_BIN), "-c", arc2, src, cfg], check=True, capture_output=True, env=ENV)
            v.update(config_hex=open(cfg, "rb").read().hex(), archive_with_config_bytes=os.path.getsize(arc2),
                     archive_with_config_sha256=sha_file(arc2))
    return v


def main():
    assert oracle.have_ref(), "oracle/_ref missing: run `make -C oracle ref` where /root/reference exists"
    names = sys.argv[1:] or list(INPUTS)
    doc = json.loads(OUT.read_text()) if OUT.exists() else dict(
        generator="tests/golden/make_golden_big.py", reference="akamiru/bce v0.4 bce.cpp (unmodified)", vectors={})
    for name in names:
        t0 = time.time()
        doc["vectors"][name] = one(name)
        OUT.write_text(json.dumps(doc, indent=1))
        print(name, doc["vectors"][name]["archive_bytes"], f"{time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main()
