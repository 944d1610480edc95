"""Seeded inputs shared by the CPU and GPU parity tests (the edge cases SURVEY.md 4 lists)."""
from __future__ import annotations

import numpy as np

from bce_b200 import synth


def rnd(n, sigma, seed):
    return np.random.default_rng(seed).integers(0, sigma, size=n, dtype=np.uint8).tobytes()


def small_cases():
    """(name, bytes, primitive?) -- small enough for every oracle path."""
    cases = [
        ("one-byte", b"x", True),
        ("two-bytes", b"ba", True),
        ("kat-hello", b"hello world, hello world! the quick brown fox jumps over the lazy dog", True),
        ("kat-run", b"a" * 999 + b"b", True),
        ("abab-power", b"abab", False),
        ("babbab-power", b"babbab", False),
        ("cabx3-power", b"cabcabcab", False),
        ("all-same", b"z" * 257, False),
        ("block-x16", rnd(64, 4, 7) * 16, False),
        ("len7", b"mississ", True),
        ("len8", b"abcdefgh", True),
        ("len9", b"ippississ", True),
        ("len31", rnd(31, 3, 1), True),
        ("len32", rnd(32, 3, 2), True),
        ("len33", rnd(33, 3, 3), True),
        ("binary-2sym", rnd(5000, 2, 4), True),
        ("dna-4sym", rnd(20000, 4, 5), True),
        ("bytes-256", rnd(30000, 256, 6), True),
        ("zero-runs", (b"\0" * 700 + b"\1\2\3") * 3 + b"\7", True),
        ("tile-edge-4096", rnd(4096, 16, 8), True),
        ("tile-edge-4097", rnd(4097, 16, 9), True),
        ("tile-edge-3072", rnd(3072, 200, 10), True),
        ("tile-edge-1024", rnd(1024, 5, 11), True),
        ("long-repeat", rnd(3000, 50, 12) + rnd(1000, 50, 13) + rnd(3000, 50, 12) + b"!", True),
    ]
    return cases


def medium_cases():
    return [
        ("markov2-200k", synth.generate("markov2-text", 200_000, 1).tobytes(), True),
        ("enwik-300k", synth.generate("enwik-shaped", 300_000, 2).tobytes(), True),
        ("mixed-2MiB+5", synth.generate("mixed-binary", (2 << 20) + 5, 4).tobytes(), True),
        ("uniform-100k", synth.generate("uniform", 100_000, 9).tobytes(), True),
    ]
