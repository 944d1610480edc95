"""GPU: BASELINE.json's full-size configurations through size-independent properties (the CPU
oracle would need minutes there).  What is checked needs no second implementation:

  * SA is a permutation, offset = SA[0], L[r] = T[(SA[r]-1) mod n], sampled neighbours are in
    rotation order, the BWT has the byte histogram of T;
  * the rank words are self-consistent (cumulative rank = prefix popcount) and
    inverse(wavelet(BWT)) == T  -- encode -> decode round trip on the device;
  * the level loop visits exactly n-1 nodes per level and the whole archive
    written from the device's packed counts decodes to the input with the UNMODIFIED reference
    decoder (`bce_ref -d`), at a size where that takes seconds."""
import numpy as np
import pytest

from bce_b200 import host, synth
from bce_b200.gpu import EMIT_CODER
from oracle import oracle

pytestmark = pytest.mark.gpu


def check_bwt_properties(fe, T):
    n = T.size
    L, off, sa = fe.bwt(T, want_sa=True)
    assert off == int(sa[0])
    seen = np.zeros(n, dtype=np.uint8)
    seen[sa] = 1
    assert int(seen.sum()) == n                                   # permutation
    assert (L == T[(sa.astype(np.int64) - 1) % n]).all()
    assert (np.bincount(L, minlength=256) == np.bincount(T, minlength=256)).all()
    rng = np.random.default_rng(1)
    TT = np.concatenate([T, T[:4096]])
    for r in rng.integers(0, n - 1, size=3000):
        a, b = int(sa[r]), int(sa[r + 1])
        x, y = TT[a:a + 4096].tobytes(), TT[b:b + 4096].tobytes()
        assert x <= y, (r, a, b)
    return L, off


def check_wavelet_and_inverse(fe, T, L, off):
    n = T.size
    ranks, Cv = fe.wavelet(L)
    for j in (0, 3, 7):
        bits = (ranks[j] >> np.uint64(32)).astype(np.uint32)
        pop = np.zeros(bits.size, dtype=np.uint64)
        v = bits.copy()
        for _ in range(32):
            pop += (v & 1).astype(np.uint64)
            v >>= 1
        cum = np.concatenate([[0], np.cumsum(pop)[:-1]]).astype(np.uint64)
        assert ((ranks[j] & np.uint64(0xFFFFFFFF)) == cum).all()
    ones0 = int(ranks[0][n // 32] & np.uint64(0xFFFFFFFF)) + bin(int(ranks[0][n // 32] >> np.uint64(32)) & ((1 << (n % 32)) - 1)).count("1")
    assert ones0 == int((L & 1).sum())
    assert Cv[1] == n - ones0                                     # C[i] = zeros of level (i+7)%8
    back = fe.unbwt(ranks, off, n)
    assert (back == T).all()
    st = fe.stats()
    return st


def check_level_loop(fe, T):
    n = T.size
    off, Cv, words = fe.compress_front_words(T, EMIT_CODER)
    st = fe.stats()
    assert st["cse_visits"] == 8 * (n - 1)
    assert st["cse_words"] == sum(w.size for w in words)
    return off, Cv, words


@pytest.mark.parametrize("workload", ["enwik-100MB", "mixed-256MB"])
def test_fullsize_properties(frontend, workload):
    kind, n, seed = synth.CONFIGS[workload]
    T = synth.generate(kind, n, seed)
    L, off = check_bwt_properties(frontend, T)
    check_wavelet_and_inverse(frontend, T, L, off)
    check_level_loop(frontend, T)


def test_archive_decodes_with_the_unmodified_reference(frontend):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not present")
    T = synth.generate("enwik-shaped", 12_000_000, 21)
    arc = host.compress(frontend, T, threads=8)
    assert oracle.ref_decompress(arc) == T.tobytes()


def test_scan_then_compress_with_config(frontend):
    """configs[3]: bce -s then bce -c ... cfg on mixed binary / low-entropy data (reduced size here;
    the config is checked against the host collector fed with raw counts)."""
    T = synth.generate("mixed-binary", (8 << 20) + 3, 4)
    cfg = host.scan(frontend, T)
    off, Cv, streams = frontend.compress_front(T)
    assert cfg == host.scan_config(streams)
    arc = host.compress(frontend, T, cfg=cfg, threads=8)
    assert arc == host.encode_archive(Cv, streams, T.size, off, cfg=cfg, threads=8)
    if oracle.have_ref():
        assert oracle.ref_decompress(arc) == T.tobytes()


def test_two_gigabyte_input_near_the_size_limit(frontend):
    """n = 2.1e9 (the format allows n <= 2^31 - 1): every 32-bit position, tile index and descriptor
    offset is exercised close to its limit.  Checked by the device round trip
    inverse(wavelet(BWT(T))) == T and the level loop's visit count 8(n - 1)."""
    n = 2_100_000_000
    T = np.frombuffer(synth.generate("enwik-shaped", n, 9), dtype=np.uint8)
    L, off, _ = frontend.bwt(T)
    assert (np.bincount(L, minlength=256) == np.bincount(T, minlength=256)).all()
    ranks, _ = frontend.wavelet(L)
    out = np.frombuffer(frontend.unbwt(ranks, off, n), dtype=np.uint8)
    assert np.array_equal(out, T)
    del ranks, L, out
    frontend.stage_input(T)
    frontend.front_resident()
    st = frontend.stats()
    assert st["cse_visits"] == 8 * (n - 1)
