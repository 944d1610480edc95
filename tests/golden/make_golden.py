"""Generates tests/golden/ref_vectors.json from the UNMODIFIED reference (oracle/_ref, built by
oracle/Makefile from /root/reference/bce.cpp).  Run in the build container only:

    python tests/golden/make_golden.py

Each vector pins what the reference itself produced for a seeded input: BWT, offset, C[8], the
(s,k,c1,c2,cs) call sequence per stream (count + sha256), and the archive (`bce_ref -c`).
kat.json holds the three whole-archive known answers listed in SURVEY.md 4."""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import oracle  # noqa: E402
from tests.inputs import medium_cases, small_cases  # noqa: E402


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def main():
    assert oracle.have_ref(), "oracle/_ref missing: run `make -C oracle ref` where /root/reference exists"
    vectors = []
    for name, data, primitive in small_cases() + medium_cases():
        r = oracle.ref_front(data, want_bwt=True, want_ranks=True)
        arc = oracle.ref_compress(data)
        v = dict(name=name, n=len(data), primitive=primitive, input_sha256=sha(data),
                 offset=r["offset"], bwt_sha256=sha(r["bwt"].tobytes()), C=r["C"],
                 rank_sha256=[sha(x.tobytes()) for x in r["ranks"]],
                 stream_counts=[int(s.shape[0]) for s in r["streams"]],
                 stream_sha256=[sha(s.tobytes()) for s in r["streams"]],
                 archive_bytes=len(arc), archive_sha256=sha(arc))
        if primitive:
            assert oracle.ref_decompress(arc) == data, name
        vectors.append(v)
        print(name, len(data), len(arc))
    out = ROOT / "tests" / "golden" / "ref_vectors.json"
    out.write_text(json.dumps(dict(generator="tests/golden/make_golden.py", reference="akamiru/bce v0.4 bce.cpp (unmodified)",
                                   vectors=vectors), indent=1))
    # -s: config produced by the reference for one input
    import subprocess, tempfile, os
    data = dict((c[0], c[1]) for c in medium_cases())["markov2-200k"]
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "in"), "wb").write(data)
        subprocess.run([str(oracle.REF_BIN), "-s", os.path.join(d, "cfg"), os.path.join(d, "in")], check=True, capture_output=True)
        cfg = open(os.path.join(d, "cfg"), "rb").read()
        arc = oracle.ref_compress(data, cfg_path=os.path.join(d, "cfg"))
    (ROOT / "tests" / "golden" / "scan_markov2_200k.json").write_text(json.dumps(
        dict(input="markov2-200k", config_hex=cfg.hex(), archive_with_config_sha256=sha(arc), archive_bytes=len(arc))))


if __name__ == "__main__":
    main()
