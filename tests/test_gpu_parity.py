"""GPU parity tests proper: every stage of the CUDA path, called through the C ABI
(include/bce_gpu.h via bce_b200.gpu), against the CPU oracle on the same seeded inputs.
Bit-exact is the bar: integers, bytes and indices only on this path."""
import numpy as np
import pytest

from bce_b200 import host
from oracle import oracle
from tests.inputs import medium_cases, small_cases

pytestmark = pytest.mark.gpu

CASES = small_cases() + medium_cases()
IDS = [c[0] for c in CASES]


def first_diff(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        return f"shape {a.shape} vs {b.shape}"
    d = np.nonzero(a.reshape(-1) != b.reshape(-1))[0]
    if d.size == 0:
        return None
    i = int(d[0])
    return f"{d.size} mismatches, first at flat index {i}: got {a.reshape(-1)[i]} want {b.reshape(-1)[i]}"


@pytest.mark.parametrize("name,data,primitive", CASES, ids=IDS)
def test_bwt_offset_sa(frontend, name, data, primitive):
    L, off, sa = frontend.bwt(data, want_sa=True)
    Lo, offo, sao = oracle.bwt(data, want_sa=True)
    assert off == offo
    assert first_diff(L, Lo) is None, first_diff(L, Lo)
    assert first_diff(sa, sao) is None, first_diff(sa, sao)
    st = frontend.stats()
    assert st["sort_rounds"] >= 1 and st["sort_m"][0] == len(data)


@pytest.mark.parametrize("name,data,primitive", CASES, ids=IDS)
def test_wavelet_ranks(frontend, name, data, primitive):
    Lo, _, _ = oracle.bwt(data)
    ranks, Cv = frontend.wavelet(Lo)
    ro = oracle.wavelet(Lo)
    for j in range(8):
        assert first_diff(ranks[j], ro[j]) is None, (j, first_diff(ranks[j], ro[j]))
    n = len(data)
    zeros = [n - (int(r[n // 32]) & 0xFFFFFFFF) - bin((int(r[n // 32]) >> 32) & ((1 << (n % 32)) - 1)).count("1") for r in ro]
    assert Cv == [zeros[(i + 7) % 8] for i in range(8)]


@pytest.mark.parametrize("name,data,primitive", CASES, ids=IDS)
def test_cse_streams(frontend, name, data, primitive):
    Lo, _, _ = oracle.bwt(data)
    want = oracle.cse(oracle.wavelet(Lo), len(data))
    Cv, streams = frontend.cse(Lo)
    assert Cv == want["C"]
    for i in range(8):
        assert first_diff(streams[i], want["streams"][i]) is None, (i, first_diff(streams[i], want["streams"][i]))
    st = frontend.stats()
    assert st["cse_visits"] == sum(want["visits"])
    assert st["cse_rounds"] == want["rounds"]
    if primitive and len(data) > 1:
        assert st["cse_visits"] == 8 * (len(data) - 1)


@pytest.mark.parametrize("name,data,primitive", CASES, ids=IDS)
def test_fused_front_archive(frontend, name, data, primitive):
    """fused call -> counts -> archive writer == the oracle's `bce -c` archive (KATs included)."""
    off, Cv, streams = frontend.compress_front(data)
    got = oracle.encode_archive(Cv, streams, len(data), off)
    assert got == oracle.compress(data)


@pytest.mark.parametrize("name,data,primitive", CASES, ids=IDS)
def test_unbwt(frontend, name, data, primitive):
    if not primitive:
        pytest.skip("the reference does not round-trip exact powers (SURVEY.md 4-6)")
    Lo, off, _ = oracle.bwt(data)
    ranks = oracle.wavelet(Lo)
    out = frontend.unbwt(ranks, off, len(data))
    assert out.tobytes() == data
    if len(data) <= 40000:
        assert first_diff(out, oracle.unbwt_bitwise(ranks, off, len(data))) is None


def test_small_emission_buffers_force_draining(frontend):
    """Several cse_next batches must concatenate to the same streams."""
    from bce_b200 import synth
    data = synth.generate("markov2-text", 400_000, 3).tobytes()
    Lo, _, _ = oracle.bwt(data)
    want = oracle.cse(oracle.wavelet(Lo), len(data))
    from bce_b200.gpu import OPT_EMIT_BATCH_BYTES
    frontend.set_option(OPT_EMIT_BATCH_BYTES, 8 * 20 * 210_000)
    try:
        Cv, streams = frontend.cse(Lo)
        batches = frontend.last_batches
    finally:
        frontend.set_option(OPT_EMIT_BATCH_BYTES, 0)
    assert batches > 1, "the option did not force several batches"
    for i in range(8):
        assert first_diff(streams[i], want["streams"][i]) is None, i


def test_resident_matches_hosted(frontend):
    """The device-resident run (bench.py's `value` leg) emits, word for word and in order, what the hosted run
    hands back -- in every emission mode."""
    from bce_b200 import synth
    from bce_b200.gpu import EMIT_CODER, EMIT_RAW, EMIT_SCAN, OPT_RESIDENT_CHECKSUM
    data = synth.generate("enwik-shaped", 500_000, 5).tobytes()
    off, Cv, streams = frontend.compress_front(data)
    frontend.set_option(OPT_RESIDENT_CHECKSUM, 1)
    try:
        for mode in (EMIT_RAW, EMIT_CODER, EMIT_SCAN):
            if mode == EMIT_RAW:
                hosted = [s.reshape(-1) for s in streams]
            else:
                _, _, hosted = frontend.compress_front_words(data, mode)
            frontend.set_emit_mode(mode)
            frontend.stage_input(data)
            off2, total = frontend.front_resident()
            got = frontend.resident_checksum()
            frontend.set_emit_mode(EMIT_RAW)
            assert off2 == off
            assert total == (sum(s.shape[0] for s in streams) if mode == EMIT_RAW else sum(w.size for w in hosted))
            assert got == [oracle.word_checksum(w) for w in hosted], mode
    finally:
        frontend.set_emit_mode(EMIT_RAW)
        frontend.set_option(OPT_RESIDENT_CHECKSUM, 0)


def test_prefetched_input_is_the_input(frontend):
    """bce_gpu_prefetch_input: the next input's upload started while the current level loop is still being drained.
    A matching call finds the text on the device, any other input is uploaded as usual; results never depend on it."""
    import ctypes as C

    from bce_b200 import synth
    from bce_b200.gpu import CseWords, EMIT_CODER
    lib, h = frontend.lib, frontend.h

    def pinned(data):
        p = lib.bce_gpu_host_alloc(h, len(data))
        assert p
        C.memmove(p, data, len(data))
        return p

    A = synth.generate("enwik-shaped", 300_000, 11).tobytes()
    B = synth.generate("markov2-text", 200_000, 12).tobytes()
    D = synth.generate("mixed-binary", 150_000, 13).tobytes()
    pa, pb, pd = pinned(A), pinned(B), pinned(D)
    try:
        def run(ptr, n, prefetch=None):
            frontend.set_emit_mode(EMIT_CODER)
            off, Cv = C.c_uint32(), (C.c_uint32 * 8)()
            frontend._check(lib.bce_gpu_compress_front(h, ptr, n, C.byref(off), Cv))
            if prefetch:
                frontend._check(lib.bce_gpu_prefetch_input(h, prefetch[0], prefetch[1]))
            words = [[] for _ in range(8)]
            b = CseWords()
            while True:
                frontend._check(lib.bce_gpu_cse_next_words(h, C.byref(b)))
                for i in range(8):
                    if b.count[i]:
                        words[i].append(np.ctypeslib.as_array((C.c_uint32 * int(b.count[i])).from_address(C.addressof(b.words[i].contents))).copy())
                if b.done:
                    break
            frontend.set_emit_mode(0)
            return int(off.value), [int(x) for x in Cv], [np.concatenate(w) if w else np.zeros(0, np.uint32) for w in words]

        want = {k: frontend.compress_front_words(v, EMIT_CODER) for k, v in (("A", A), ("B", B), ("D", D))}

        def same(got, key):
            assert got[0] == want[key][0] and got[1] == want[key][1], key
            for i in range(8):
                assert first_diff(got[2][i], want[key][2][i]) is None, (key, i)

        same(run(pa, len(A), prefetch=(pb, len(B))), "A")      # B travels while A's counts are fetched
        same(run(pb, len(B), prefetch=(pa, len(A))), "B")      # ... and is found on the device; A is prefetched
        same(run(pd, len(D)), "D")                             # a different input than the one prefetched: uploaded as usual
        same(run(pa, len(A), prefetch=(pa, len(A))), "A")
        same(run(pa, len(A)), "A")                             # the same buffer again, prefetched by the call before
        assert lib.bce_gpu_prefetch_input(h, D, len(D)) == 0   # pageable memory: accepted, nothing happens
        same(run(pd, len(D)), "D")
    finally:
        for p in (pa, pb, pd):
            lib.bce_gpu_host_free(h, p)


def test_twenty_bit_words_are_the_same_words(frontend):
    """bce_gpu_cse_next_words20: the CODER words of a batch as 20 bits each (packed on the device before the copy),
    on inputs with many k > 31 counts (three-word escapes, k field 0) and over several batches."""
    from bce_b200 import synth
    from bce_b200.gpu import EMIT_CODER, OPT_EMIT_BATCH_BYTES
    cases = [synth.generate("mixed-binary", (1 << 20) + 77, 4).tobytes(), synth.generate("enwik-shaped", 400_000, 9).tobytes(),
             b"a" * 999 + b"b", b"hello world, hello world! the quick brown fox jumps over the lazy dog"]
    seen_escape = False
    for data in cases:
        _, Cv, words = frontend.compress_front_words(data, EMIT_CODER)
        seen_escape |= any((((w >> 5) & 31) == 0).any() for w in words if w.size)
        frontend.set_option(OPT_EMIT_BATCH_BYTES, 1 << 20)
        try:
            _, Cv2, words20 = frontend.compress_front_words20(data)
        finally:
            frontend.set_option(OPT_EMIT_BATCH_BYTES, 0)
        assert Cv2 == Cv
        for i in range(8):
            assert first_diff(words20[i], words[i]) is None, i
            assert not words[i].size or int(words[i].max()) < (1 << 20)
    assert seen_escape, "no k > 31 count in any case"


def test_slot_layout_forced_on_small_inputs(frontend):
    """Frontiers of millions of nodes per round live in the slot layout (csrc/cse_slots.cuh): per-chunk output slots, a
    scan of the slot counts between rounds, conversion from and back to the flat layout.  That only switches on at
    sizes the oracle cannot check, so it is forced here on everything from 16 K bytes up: every count of every stream
    (raw), the packed words' archive, and several small batches (the kernel stops to drain while in slots)."""
    from bce_b200.gpu import OPT_EMIT_BATCH_BYTES, OPT_SLOT_ENTER_NODES
    frontend.set_option(OPT_SLOT_ENTER_NODES, 2048)
    try:
        used = 0
        for name, data, _ in small_cases() + medium_cases():
            if len(data) < 8 * 2048:
                continue
            Lo, offo, _ = oracle.bwt(data)
            want = oracle.cse(oracle.wavelet(Lo), len(data))
            off, Cv, streams = frontend.compress_front(data)
            st = frontend.stats()
            assert off == offo and Cv == want["C"], name
            for i in range(8):
                assert first_diff(streams[i], want["streams"][i]) is None, (name, i)
            assert st["cse_visits"] == sum(want["visits"]) and st["cse_rounds"] == want["rounds"], name
            assert host.compress(frontend, data, threads=1) == oracle.compress(data), name
            frontend.set_option(OPT_EMIT_BATCH_BYTES, 1 << 20)
            try:
                _, streams2 = frontend.cse(Lo)
                assert frontend.last_batches > 1
            finally:
                frontend.set_option(OPT_EMIT_BATCH_BYTES, 0)
            for i in range(8):
                assert first_diff(streams2[i], want["streams"][i]) is None, (name, i, "batched")
            used += 1
        assert used >= 5
    finally:
        frontend.set_option(OPT_SLOT_ENTER_NODES, 0)


def test_mid_kernel_sees_every_frontier_size(frontend):
    """Frontiers between the cluster kernels' few thousand nodes and the wide kernel's hundreds of thousands run in
    cse_mid_kernel (csrc/cse_mid.cuh: slot layout, one grid barrier per round, every CTA scans the slot directory
    itself, flat layout read on entry and written back on exit).  With the cluster / one-CTA kernels switched off it
    runs every round from the roots on: every count of every stream (raw), rounds and visits, the packed words'
    archive, and small batches (the kernel leaves to drain and comes back from the flat layout)."""
    from bce_b200.gpu import OPT_EMIT_BATCH_BYTES, OPT_NO_NARROW_KERNELS, OPT_MID_ENTER_NODES
    frontend.set_option(OPT_NO_NARROW_KERNELS, 1)
    try:
        used = 0
        for name, data, _ in small_cases() + medium_cases():
            if len(data) < 16384:
                continue
            Lo, offo, _ = oracle.bwt(data)
            want = oracle.cse(oracle.wavelet(Lo), len(data))
            for enter in (0, 3000):                 # default hand-over to the wide kernel, and a very early one (ping-pong)
                frontend.set_option(OPT_MID_ENTER_NODES, enter)
                off, Cv, streams = frontend.compress_front(data)
                st = frontend.stats()
                assert off == offo and Cv == want["C"], name
                for i in range(8):
                    assert first_diff(streams[i], want["streams"][i]) is None, (name, i, enter)
                assert st["cse_visits"] == sum(want["visits"]) and st["cse_rounds"] == want["rounds"], (name, enter)
                # small batches: the kernel leaves to drain and is entered again from the flat layout at every frontier
                # size (with the early hand-over its instances for 32- / 64- / 128-node chunks all get their turn)
                frontend.set_option(OPT_EMIT_BATCH_BYTES, 1 << 18)
                try:
                    _, streams2 = frontend.cse(Lo)
                    assert frontend.last_batches > 1
                finally:
                    frontend.set_option(OPT_EMIT_BATCH_BYTES, 0)
                for i in range(8):
                    assert first_diff(streams2[i], want["streams"][i]) is None, (name, i, enter, "batched")
            frontend.set_option(OPT_MID_ENTER_NODES, 0)
            assert host.compress(frontend, data, threads=1) == oracle.compress(data), name
            used += 1
        assert used >= 5
    finally:
        frontend.set_option(OPT_NO_NARROW_KERNELS, 0)
        frontend.set_option(OPT_MID_ENTER_NODES, 0)


def test_scan_buckets_from_the_device(frontend):
    """bce_gpu_cse_next_buckets: per stream and (k, key) the device's runs of symbol bytes, joined over
    the batches, are the stream's SCAN words of that bucket in order; keys first appear in stream order; the
    halvings add up.  Small batches force several of them."""
    from bce_b200 import synth
    from bce_b200.gpu import EMIT_SCAN, OPT_EMIT_BATCH_BYTES
    for kind, n, seed in (("markov2-text", 300_000, 3), ("mixed-binary", (1 << 20) + 77, 4)):
        data = synth.generate(kind, n, seed).tobytes()
        _, _, words = frontend.compress_front_words(data, EMIT_SCAN)
        frontend.set_option(OPT_EMIT_BATCH_BYTES, 4 << 20)
        try:
            _, _, batches = frontend.compress_front_buckets(data)
        finally:
            frontend.set_option(OPT_EMIT_BATCH_BYTES, 0)
        assert len(batches) > 1
        for i in range(8):
            want_syms, want_bk, want_halv = host.bucket_scan_words(words[i])
            want = {}
            ends = list(want_bk["start"][1:]) + [want_syms.size]
            for b, e in zip(want_bk, ends):
                want[int(b["key"])] = want_syms[int(b["start"]):int(e)].tobytes()
            order_want = [int(k) for k in want_bk["key"][np.argsort(want_bk["first"], kind="stable")]]
            got, order_got, halv, seen = {}, [], 0, 0
            for one in batches:
                syms, bk, h = one[i]
                halv += h
                bs = bk[np.argsort(bk["start"], kind="stable")]
                ends = list(bs["start"][1:]) + [syms.size]
                for b, e in zip(bs, ends):
                    got[int(b["key"])] = got.get(int(b["key"]), b"") + syms[int(b["start"]):int(e)].tobytes()
                for b in bk[np.argsort(bk["first"], kind="stable")]:
                    if int(b["key"]) not in order_got:
                        order_got.append(int(b["key"]))
                seen += syms.size
            assert seen == words[i].size and halv == want_halv, i
            assert got == want, i
            assert order_got == order_want, i


PACKED = [c for c in CASES if c[0] in ("kat-hello", "kat-run", "bytes-256", "long-repeat", "markov2-200k", "mixed-2MiB+5")]


@pytest.mark.parametrize("name,data,primitive", PACKED, ids=[c[0] for c in PACKED])
def test_packed_emission_matches_host_packer(frontend, name, data, primitive):
    """BCE_EMIT_CODER / BCE_EMIT_SCAN words from the device == the host packer applied to the
    oracle's raw counts, word for word."""
    from bce_b200 import host
    from bce_b200.gpu import EMIT_CODER, EMIT_SCAN
    Lo, offo, _ = oracle.bwt(data)
    want = oracle.cse(oracle.wavelet(Lo), len(data))
    for mode in (EMIT_CODER, EMIT_SCAN):
        off, Cv, words = frontend.compress_front_words(data, mode)
        ref = host.pack_counts(mode, want["streams"])
        assert off == offo and Cv == want["C"]
        for i in range(8):
            assert first_diff(words[i], ref[i]) is None, (mode, i, first_diff(words[i], ref[i]))


@pytest.mark.parametrize("name,data,primitive", PACKED, ids=[c[0] for c in PACKED])
def test_compress_and_scan_pipelines(frontend, name, data, primitive):
    """bce -c / bce -s of this repository: GPU front end (packed words) + host coders."""
    from bce_b200 import host
    assert host.compress(frontend, data, threads=8) == oracle.compress(data)
    want = oracle.cse(oracle.wavelet(oracle.bwt(data)[0]), len(data))
    assert host.scan(frontend, data) == host.scan_config(want["streams"])


@pytest.mark.parametrize("name,data,primitive", [c for c in CASES if c[2]], ids=[c[0] for c in CASES if c[2]])
def test_decompress_with_gpu_inverse(frontend, name, data, primitive):
    """`bce -d`: host decoder + bce_gpu_unbwt, on archives written by this repository's bce -c path."""
    from bce_b200 import host
    arc = host.compress(frontend, data, threads=4)
    assert arc == oracle.compress(data)
    assert host.decompress(arc) == data


def test_random_small_inputs_archive_parity(frontend):
    """Property check over many seeded small inputs (sizes around the tile and word edges, tiny
    and large alphabets, planted repeats): archive from the GPU front end == oracle archive."""
    from bce_b200 import host
    rng = np.random.default_rng(2024)
    for trial in range(120):
        n = int(rng.choice([1, 2, 3, 5, 8, 9, 31, 32, 33, 63, 64, 65, 255, 256, 257, 511, 1023, 1024, 1025,
                            2047, 2049, 4095, 4096, 4097, 5000, 12289]))
        sigma = int(rng.choice([1, 2, 3, 4, 16, 64, 256]))
        data = rng.integers(0, sigma, size=n, dtype=np.uint8)
        if n > 64 and trial % 3 == 0:                      # plant a long repeat
            k = int(rng.integers(8, n // 2))
            src = int(rng.integers(0, n - k))
            dst = int(rng.integers(0, n - k))
            data[dst:dst + k] = data[src:src + k].copy()
        data = data.tobytes()
        got = host.compress(frontend, data, threads=1)
        assert got == oracle.compress(data), (trial, n, sigma)


def test_tile_local_sort_forced_on_small_inputs(frontend):
    """Rounds >= 1 of the suffix sort normally go tile by tile (csrc/local_sort.cuh) only for working sets
    of 2^20 rotations and more, i.e. never on the inputs the oracle can check.  Forced on everything here:
    groups crossing tile boundaries (fall-back radix sort + placement), whole-tile groups (bail-out to the
    radix path), partial last tiles, the index tie-break round of exact powers."""
    from bce_b200.gpu import OPT_LOCAL_SORT_MIN
    frontend.set_option(OPT_LOCAL_SORT_MIN, 1)
    local = 0
    try:
        for name, data, _ in small_cases() + medium_cases():
            n = len(data)
            if n < 2:
                continue
            Lo, offo, sao = oracle.bwt(data, want_sa=True)
            L, off, sa = frontend.bwt(data, want_sa=True)
            assert off == offo, name
            assert np.array_equal(np.asarray(sa), np.asarray(sao)), name
            assert bytes(L) == bytes(Lo), name
            local += frontend.stats()["sort_local_elems"]
    finally:
        frontend.set_option(OPT_LOCAL_SORT_MIN, 0)
    assert local > 0, "the option did not force the tile-local sort"


def test_error_paths_of_the_c_abi(frontend):
    """Call-sequence and resource errors come back as codes with a message, never as a crash or a wrong result:
    BCE_GPU_E_STATE for calls out of order, BCE_GPU_E_ARG for bad arguments, BCE_GPU_E_NOMEM / BCE_GPU_E_FRONTIER
    when the scratch limit leaves no room for the level loop's frontier."""
    import ctypes as C

    from bce_b200 import Frontend, synth
    from bce_b200.gpu import CseBatch, CseWords, EMIT_CODER, EMIT_SCAN, ScanBuckets
    fe = Frontend(0)                                   # a second context on the same device: contexts share nothing
    try:
        lib, h = fe.lib, fe.h
        assert lib.bce_gpu_cse_next(h, C.byref(CseBatch())) == -4            # no cse_begin yet
        assert lib.bce_gpu_cse_next_words(h, C.byref(CseWords())) == -4
        assert lib.bce_gpu_cse_next_buckets(h, C.byref(ScanBuckets())) == -4
        assert lib.bce_gpu_cse_begin(h, None, 1000, (C.c_uint32 * 8)()) == -4  # no BWT resident
        off, tup = C.c_uint32(), C.c_uint64()
        assert lib.bce_gpu_front_resident(h, C.byref(off), C.byref(tup)) == -4  # no staged input
        assert lib.bce_gpu_resident_checksum(h, (C.c_uint64 * 8)(), (C.c_uint64 * 8)()) == -4
        assert lib.bce_gpu_bwt(h, None, 10, None, None, None) == -1
        buf = (C.c_uint8 * 16)()
        assert lib.bce_gpu_bwt(h, buf, 0, None, None, None) == -1              # n = 0 (the reference crashes)
        assert lib.bce_gpu_bwt(h, buf, 0x80000000, None, None, None) == -1     # n >= 2^31
        assert lib.bce_gpu_set_option(h, 99, 1) == -1
        assert lib.bce_gpu_set_emit_mode(h, 7, None) == -1
        bad_cfg = (C.c_uint8 * 288)(*([9] * 288))
        assert lib.bce_gpu_set_emit_mode(h, EMIT_CODER, bad_cfg) == -1
        assert b"context bits" in lib.bce_gpu_last_error(h)

        data = synth.generate("enwik-shaped", 2_000_000, 3)
        # raw counts asked for while the run emits packed words
        fe.set_emit_mode(EMIT_CODER)
        offv, Cv = C.c_uint32(), (C.c_uint32 * 8)()
        assert lib.bce_gpu_compress_front(h, data.ctypes.data, data.size, C.byref(offv), Cv) == 0
        assert lib.bce_gpu_cse_next(h, C.byref(CseBatch())) == -4
        assert lib.bce_gpu_cse_next_buckets(h, C.byref(ScanBuckets())) == -4   # needs BCE_EMIT_SCAN
        fe.set_emit_mode(0)

        # a scratch limit that leaves the level loop's frontier no room (the suffix sort takes what it needs)
        L, off2, _ = fe.bwt(data)
        want_L, want_off, _ = oracle.bwt(data)
        assert off2 == want_off and bytes(L) == bytes(want_L)                   # the context works again after the errors
        fe.set_scratch_limit(3 << 20)
        rc = lib.bce_gpu_cse_begin(h, None, data.size, Cv)
        if rc == 0:                                                             # the frontier budget is found out while running
            b = CseBatch()
            while rc == 0 and not b.done:
                rc = lib.bce_gpu_cse_next(h, C.byref(b))
        assert rc in (-2, -5), rc
        assert len(lib.bce_gpu_last_error(h)) > 0
        fe.set_scratch_limit(0)
        assert host.compress(fe, data, threads=2) == host.compress(frontend, data, threads=2)
    finally:
        fe.close()
