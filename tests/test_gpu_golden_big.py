"""GPU: BASELINE.json's configurations at their FULL sizes, bit-exact against what the unmodified
reference produced (tests/golden/big_vectors.json, written once by tests/golden/make_golden_big.py
from oracle/_ref = /root/reference/bce.cpp compiled as it lies).  Per input:

  * the input regenerated from its seed has the recorded sha256 (the generator did not drift);
  * BWT bytes (sha256) and offset                         File::rotate + File::bwt, bce.cpp:858-910
  * C[8], the number of coder calls of every stream and their order-sensitive checksums
    (oracle/ref_tap.cpp) over the RAW counts, batch by batch       BCE::code, bce.cpp:1236-1373
  * the archive `bce -c` writes (size and sha256)                  main, bce.cpp:1403-1427
  * the device-resident run (bench.py's `value` leg) emits word for word what the hosted run hands
    back: bce_gpu_resident_checksum == checksum of the hosted CODER words
  * mixed-256MB: the 288-byte config of `bce -s` and the archive of `bce -c archive file cfg`.

Everything goes through the C ABI (ctypes); the 1 GB case covers the paths that only switch on at
scale (binned rank scatter, 1024-node level-loop tiles, two-set emission with the tail cut)."""
import hashlib
import json
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np
import pytest

from bce_b200 import host, synth
from bce_b200.gpu import EMIT_CODER, EMIT_RAW, OPT_RESIDENT_CHECKSUM
from oracle import oracle

pytestmark = pytest.mark.gpu

VECTORS = json.loads((Path(__file__).parent / "golden" / "big_vectors.json").read_text())["vectors"]
M64 = (1 << 64) - 1


def sha(b):
    return hashlib.sha256(b).hexdigest()


def stream_checksums(frontend, T, mode, fn):
    """Drain the fused front end batch by batch; fn(view, first_index) -> (sum, wsum) per stream and batch."""
    head = None
    count = [0] * 8
    acc = [(0, 0)] * 8
    with ThreadPoolExecutor(8) as pool:
        for item in frontend.iter_front_batches(T, mode):
            if item[0] == "head":
                head = item[1:]
                continue
            views = item[1]
            per = 5 if mode == EMIT_RAW else 1
            parts = list(pool.map(lambda i: fn(views[i], count[i]), range(8)))
            for i in range(8):
                acc[i] = ((acc[i][0] + parts[i][0]) & M64, (acc[i][1] + parts[i][1]) & M64)
                count[i] += views[i].size // per
    return head, count, acc


def check_against_reference(frontend, name, raw_streams=True):
    v = VECTORS[name]
    T = synth.generate(v["generator"], v["n"], v["seed"])
    assert sha(T) == v["input_sha256"], "the synthetic input generator drifted"

    L, off, _ = frontend.bwt(T)
    assert off == v["offset"]
    assert sha(L) == v["bwt_sha256"], "BWT differs from the reference's"
    del L

    if raw_streams:
        (off2, Cv), calls, sums = stream_checksums(frontend, T, EMIT_RAW, oracle.call_checksum_fast)
        assert off2 == v["offset"] and Cv == v["C"]
        assert calls == v["stream_calls"]
        assert [[f"{a:016x}", f"{b:016x}"] for a, b in sums] == v["stream_checksum"], "coder calls differ from the reference's"

    arc = host.compress(frontend, T, threads=8)
    assert len(arc) == v["archive_bytes"]
    assert sha(arc) == v["archive_sha256"], "archive differs from `bce_ref -c`"

    # the resident run of bench.py's `value` leg does the same work as the hosted one
    (_, Cv), nwords, hosted = stream_checksums(frontend, T, EMIT_CODER, oracle.word_checksum)
    assert Cv == v["C"]
    frontend.set_option(OPT_RESIDENT_CHECKSUM, 1)
    frontend.set_emit_mode(EMIT_CODER)
    try:
        frontend.stage_input(T)
        off3, total = frontend.front_resident()
        resident = frontend.resident_checksum()
    finally:
        frontend.set_emit_mode(EMIT_RAW)
        frontend.set_option(OPT_RESIDENT_CHECKSUM, 0)
    assert off3 == v["offset"] and total == sum(nwords)
    assert resident == hosted, "resident emission differs from the hosted words"
    return T


@pytest.mark.parametrize("name", ["markov2-1MB", "enwik-100MB", "batch-128MB-seed100", "batch-128MB-seed163"])
def test_config_matches_reference(frontend, name):
    if name not in VECTORS:
        pytest.skip(f"{name} not in big_vectors.json")
    check_against_reference(frontend, name, raw_streams=not name.startswith("batch"))


def test_headline_1GB_matches_reference(frontend):
    """configs[2], the configuration the bench line is quoted on."""
    if "enwik-1GB" not in VECTORS:
        pytest.skip("enwik-1GB not in big_vectors.json")
    check_against_reference(frontend, "enwik-1GB")


def test_scan_config_256MB_matches_reference(frontend):
    """configs[3]: `bce -s cfg file` then `bce -c archive file cfg` on 256 MiB of mixed binary data."""
    if "mixed-256MB" not in VECTORS:
        pytest.skip("mixed-256MB not in big_vectors.json")
    v = VECTORS["mixed-256MB"]
    T = check_against_reference(frontend, "mixed-256MB")
    cfg = host.scan(frontend, T)
    assert cfg.hex() == v["config_hex"], "config differs from `bce_ref -s`"
    arc = host.compress(frontend, T, cfg=cfg, threads=8)
    assert len(arc) == v["archive_with_config_bytes"]
    assert sha(arc) == v["archive_with_config_sha256"], "archive differs from `bce_ref -c archive file cfg`"
