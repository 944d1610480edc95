"""CPU: the product's host half (libbce_host.so: range coders, archive writer, scan policy)
against the oracle's archives and the golden fixtures.  No GPU involved: the count streams
come from the oracle, exactly what the GPU path must deliver."""
import hashlib
import json
from pathlib import Path

import pytest

from bce_b200 import host
from oracle import oracle
from tests.inputs import medium_cases, small_cases

GOLD = Path(__file__).resolve().parent / "golden"
CASES = small_cases() + medium_cases()[:3]
IDS = [c[0] for c in CASES]


def streams_of(data):
    L, off, _ = oracle.bwt(data)
    c = oracle.cse(oracle.wavelet(L), len(data))
    return off, c


def test_default_config_table():
    assert host.default_config() == oracle.default_config()


@pytest.mark.parametrize("name,data,primitive", CASES, ids=IDS)
def test_archive_writer_bit_exact(name, data, primitive):
    ref = {v["name"]: v for v in json.loads((GOLD / "ref_vectors.json").read_text())["vectors"]}[name]
    off, c = streams_of(data)
    for threads, pieces in ((1, 1), (8, 1), (8, 4)):
        arc = host.encode_archive(c["C"], c["streams"], len(data), off, threads=threads, pieces=pieces)
        assert hashlib.sha256(arc).hexdigest() == ref["archive_sha256"], (threads, pieces)


def test_kat_archives_through_host_writer():
    for v in json.loads((GOLD / "kat.json").read_text())["vectors"]:
        data = eval(v["input_py"])
        off, c = streams_of(data)
        assert host.encode_archive(c["C"], c["streams"], len(data), off).hex() == v["archive_hex"]


def test_scan_config_and_archive_with_config():
    g = json.loads((GOLD / "scan_markov2_200k.json").read_text())
    data = dict((c[0], c[1]) for c in medium_cases())[g["input"]]
    off, c = streams_of(data)
    cfg = host.scan_config(c["streams"])
    assert cfg.hex() == g["config_hex"]                      # bit-exact `bce -s` output
    arc = host.encode_archive(c["C"], c["streams"], len(data), off, cfg=cfg)
    assert len(arc) == g["archive_bytes"] and hashlib.sha256(arc).hexdigest() == g["archive_with_config_sha256"]
    assert arc == oracle.encode_archive(c["C"], c["streams"], len(data), off, cfg=cfg)


def test_bucketed_scan_batches_give_the_same_config(capfd):
    """`bce -s` with the counts bucketed before they reach the host (what the device does, restated in numpy by
    host.bucket_scan_words): same 288 bytes as the reference's config and the same nine "Result size" lines as
    the collector fed count by count -- the cost sums depend on the unordered_map's iteration order."""
    from bce_b200.gpu import EMIT_SCAN
    g = json.loads((GOLD / "scan_markov2_200k.json").read_text())
    cases = dict((c[0], c[1]) for c in small_cases() + medium_cases())
    for name, pieces in ((g["input"], 1), (g["input"], 4), ("mixed-2MiB+5", 3), ("kat-run", 1)):
        data = cases[name]
        off, c = streams_of(data)
        capfd.readouterr()
        want = host.scan_config(c["streams"])
        lines_want = capfd.readouterr().out
        batches = []
        for p in range(pieces):
            part = [s[(s.shape[0] * p) // pieces:(s.shape[0] * (p + 1)) // pieces] for s in c["streams"]]
            batches.append([host.bucket_scan_words(w) for w in host.pack_counts(EMIT_SCAN, part)])
        got = host.scan_config_buckets(batches)
        lines_got = capfd.readouterr().out
        assert got == want, (name, pieces)
        assert lines_got == lines_want and lines_got.count("Result size") == 9, (name, pieces)
        if name == g["input"]:
            assert got.hex() == g["config_hex"]


def test_overlapped_coder_threads_give_the_same_archive():
    """bce_archive_begin_words / bce_archive_wait (one persistent coder thread per stream, what
    bce_compress_buffer runs under the GPU's next batch) over several batches == one serial feed."""
    from bce_b200.gpu import EMIT_CODER
    lib = host.load_library()
    cases = dict((c[0], c[1]) for c in small_cases() + medium_cases())
    for name, pieces in (("kat-hello", 1), ("markov2-200k", 5), ("long-repeat", 3), ("bytes-256", 2)):
        data = cases[name]
        off, c = streams_of(data)
        Cv = (host.C.c_uint32 * 8)(*c["C"])
        w = lib.bce_archive_begin(len(data), Cv, None)
        keep = []
        for p in range(pieces):
            part = [s[(s.shape[0] * p) // pieces:(s.shape[0] * (p + 1)) // pieces] for s in c["streams"]]
            b, k = host._words_batch(host.pack_counts(EMIT_CODER, part), done=(p == pieces - 1))
            keep.append(k)
            assert lib.bce_archive_begin_words(w, host.C.byref(b)) == 0
            assert lib.bce_archive_wait(w) == 0
        words, nw = host.C.c_void_p(), host.C.c_size_t()
        assert lib.bce_archive_finish(w, off, host.C.byref(words), host.C.byref(nw)) == 0
        arc = host.C.string_at(words.value, nw.value * 2)
        lib.bce_host_free(words)
        assert arc == oracle.compress(data), name


def test_wide_ranges_take_the_binary_decomposition():
    """k > 31 symbols (bce.cpp:507-510) and uint32 wrap in the context index (:674)."""
    import numpy as np
    rng = np.random.default_rng(3)
    streams = []
    for i in range(8):
        k = rng.integers(2, 5000, size=400, dtype=np.uint32)
        s = (rng.integers(0, 1 << 30, size=400, dtype=np.uint32) % k).astype(np.uint32)
        cs = rng.integers(2, 1 << 31, size=400, dtype=np.uint32)
        c1 = (rng.integers(0, 1 << 31, size=400, dtype=np.uint32) % cs).astype(np.uint32)
        c2 = (rng.integers(0, 1 << 31, size=400, dtype=np.uint32) % cs).astype(np.uint32)
        streams.append(np.stack([s, k, c1, c2, cs], axis=1).astype(np.uint32))
    Cv = [5] * 8
    assert host.encode_archive(Cv, streams, 1000, 3) == oracle.encode_archive(Cv, streams, 1000, 3)


def test_packed_words_give_the_same_archive_and_config():
    """The device's packed formats (include/bce_gpu.h), produced here by the host packer from
    the oracle's raw counts, must code to the same archive / scan config as the raw counts."""
    from bce_b200.gpu import EMIT_CODER, EMIT_SCAN
    g = json.loads((GOLD / "scan_markov2_200k.json").read_text())
    cases = dict((c[0], c[1]) for c in small_cases() + medium_cases())
    for name in ("kat-hello", "kat-run", "bytes-256", "markov2-200k", "long-repeat"):
        data = cases[name]
        off, c = streams_of(data)
        words = host.pack_counts(EMIT_CODER, c["streams"])
        assert sum(w.size for w in words) >= sum(s.shape[0] for s in c["streams"])
        assert host.encode_archive_words(c["C"], words, len(data), off, threads=8) == oracle.compress(data), name
    data = cases[g["input"]]
    off, c = streams_of(data)
    cfg = bytes.fromhex(g["config_hex"])
    assert host.scan_config_words(host.pack_counts(EMIT_SCAN, c["streams"])) == cfg
    words = host.pack_counts(EMIT_CODER, c["streams"], cfg=cfg)
    arc = host.encode_archive_words(c["C"], words, len(data), off, cfg=cfg)
    assert hashlib.sha256(arc).hexdigest() == g["archive_with_config_sha256"]


def test_packed_wide_ranges():
    """k > 31 (escape word + low bits) and uint32 wrap of the context index."""
    import numpy as np
    from bce_b200.gpu import EMIT_CODER
    rng = np.random.default_rng(5)
    streams = []
    for i in range(8):
        k = rng.integers(2, 1 << 20, size=300, dtype=np.uint32)
        s = (rng.integers(0, 1 << 30, size=300, dtype=np.uint32) % k).astype(np.uint32)
        cs = rng.integers(2, 1 << 31, size=300, dtype=np.uint32)
        c1 = (rng.integers(0, 1 << 31, size=300, dtype=np.uint32) % cs).astype(np.uint32)
        c2 = (rng.integers(0, 1 << 31, size=300, dtype=np.uint32) % cs).astype(np.uint32)
        streams.append(np.stack([s, k, c1, c2, cs], axis=1).astype(np.uint32))
    Cv = [7] * 8
    words = host.pack_counts(EMIT_CODER, streams)
    assert host.encode_archive_words(Cv, words, 2000, 11) == oracle.encode_archive(Cv, streams, 2000, 11)


def test_twenty_bit_words_code_the_same_archive():
    """bce_cse_words20 (two words in 5 bytes) through the overlapped writer: many k > 31 counts (three-word escapes, k
    field 0), streams longer than the unpacker's 4096-word blocks so that escapes straddle block ends, several batches."""
    import numpy as np
    from bce_b200.gpu import EMIT_CODER
    rng = np.random.default_rng(6)
    streams = []
    for i in range(8):
        m = 9000 + 501 * i
        wide = rng.random(m) < 0.3                                   # a third of the counts have k > 31
        k = np.where(wide, rng.integers(32, 1 << 20, size=m), rng.integers(2, 32, size=m)).astype(np.uint32)
        s = (rng.integers(0, 1 << 30, size=m, dtype=np.uint32) % k).astype(np.uint32)
        cs = rng.integers(2, 1 << 31, size=m, dtype=np.uint32)
        c1 = (rng.integers(0, 1 << 31, size=m, dtype=np.uint32) % cs).astype(np.uint32)
        c2 = (rng.integers(0, 1 << 31, size=m, dtype=np.uint32) % cs).astype(np.uint32)
        streams.append(np.stack([s, k, c1, c2, cs], axis=1).astype(np.uint32))
    Cv = [7] * 8
    words = host.pack_counts(EMIT_CODER, streams)
    assert all(int(w.max()) < (1 << 20) for w in words)
    want = oracle.encode_archive(Cv, streams, 2000, 11)
    assert host.encode_archive_words(Cv, words, 2000, 11) == want
    for pieces in (1, 3):
        assert host.encode_archive_words20(Cv, words, 2000, 11, pieces=pieces) == want, pieces


@pytest.mark.parametrize("name,data,primitive", [c for c in CASES if c[2]], ids=[c[0] for c in CASES if c[2]])
def test_host_decoder_low_memory_path(name, data, primitive):
    """`bce -ds`: header + 8 adaptive decoders + the level loop in decode mode + serial inverse,
    all host code (csrc/host/decode.cpp); archives come from the oracle (= the reference's)."""
    arc = oracle.compress(data)
    assert host.decompress(arc, low_memory=True) == data
    if oracle.have_ref() and len(data) <= 50000:
        assert oracle.ref_decompress(arc, low_mem=True) == data


def test_host_decoder_rejects_truncated_archives():
    arc = oracle.compress(b"hello world, hello world! the quick brown fox jumps over the lazy dog")
    with pytest.raises(RuntimeError):
        host.decompress(arc[:2], low_memory=True)


def test_damaged_archives_never_crash_the_decoder():
    """the reference's reader trusts the archive; ours must end in an error code
    (or in some output) for any input, never in an out-of-bounds access or an endless loop."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    env = dict(os.environ, BCE_HOST_MAX_N="100000")
    script = Path(__file__).resolve().parent / "fuzz_decoder.py"
    r = subprocess.run([sys.executable, str(script), "7", "1500"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout[-300:], r.stderr[-300:])
    assert "err" in r.stdout.splitlines()[-1]
