"""The `bce` command line tool (bce_b200/bce, csrc/host/bce_main.cpp) against the reference's interface,
main() bce.cpp:1376-1484: same arguments, messages and exit codes.

CPU part: usage and argument errors exit 0 (:1473-1483), a missing input exits -1 with "Error loading file"
(:1390-1393, :1412-1415), a missing archive -1 with "Archive not found." (:1435-1438), an unreadable or
damaged archive -2 with "Could not read Archive." (:1445-1448), a bad config is reported and ignored
(:629-632), `-ds` decodes the reference's own known-answer archives without a GPU.
GPU part (marked): -c / -d / -ds / -s end to end, archives and configs bit-exact against the oracle and the
golden config `bce_ref -s` wrote."""
import json
import subprocess
from pathlib import Path

import pytest

from bce_b200 import build, synth
from oracle import oracle

ROOT = Path(__file__).resolve().parent.parent
GOLD = Path(__file__).resolve().parent / "golden"
USAGE_FIRST = "Usage:"


@pytest.fixture(scope="module")
def bce():
    build.build_all()
    assert build.BIN_BCE.exists(), "the bce tool was not built"
    return str(build.BIN_BCE)


def run(bce, *args, cwd=None):
    r = subprocess.run([bce, *map(str, args)], capture_output=True, text=True, cwd=cwd, timeout=600)
    code = r.returncode if r.returncode < 128 else r.returncode - 256      # main returns -1 / -2: seen as 255 / 254
    return code, r.stdout


def test_usage_and_argument_errors_exit_zero(bce, tmp_path):
    for args in ([], ["-c"], ["-c", "a.bce"], ["-x", "a", "b"], ["c", "a", "b"], ["-d", "only-one"],
                 ["-s", "cfg"], ["-c", "a", "b", "c", "d"]):
        code, out = run(bce, *args, cwd=tmp_path)
        assert code == 0, args
        assert USAGE_FIRST in out and "bce -c archive.bce file [config.bcc]" in out, args
        assert "bce -d file archive.bce" in out and "bce -s config.bcc file" in out, args
    assert not list(tmp_path.iterdir()), "an argument error must not create files"


def test_missing_input_is_minus_one(bce, tmp_path):
    for args in (["-c", tmp_path / "a.bce", tmp_path / "nope"], ["-s", tmp_path / "cfg", tmp_path / "nope"]):
        code, out = run(bce, *args)
        assert code == -1 and "Error loading file" in out
    assert not (tmp_path / "a.bce").exists() and not (tmp_path / "cfg").exists()
    empty = tmp_path / "empty"
    empty.write_bytes(b"")
    code, out = run(bce, "-c", tmp_path / "a.bce", empty)          # the reference crashes on n = 0: refused here
    assert code == -1 and "Error loading file" in out


def test_missing_and_damaged_archives(bce, tmp_path):
    code, out = run(bce, "-d", tmp_path / "out", tmp_path / "nope.bce")
    assert code == -1 and "Archive not found." in out
    kat = json.loads((GOLD / "kat.json").read_text())["vectors"][1]
    good = bytes.fromhex(kat["archive_hex"])
    for name, blob in (("one-byte", good[:1]), ("odd", good[:-1]), ("odd+1", good + b"\x00"), ("empty", b"")):
        p = tmp_path / (name + ".bce")
        p.write_bytes(blob)
        for flag in ("-d", "-ds"):
            code, out = run(bce, flag, tmp_path / "out", p)
            assert code == -2 and "Could not read Archive." in out, (name, flag)
            assert not (tmp_path / "out").exists()
    # truncated at a word boundary: decoding fails cleanly (no crash, no output file)
    p = tmp_path / "cut.bce"
    p.write_bytes(good[: len(good) // 2 & ~1])
    code, out = run(bce, "-ds", tmp_path / "out", p)
    assert code != 0 and ("Decoding failed" in out) and not (tmp_path / "out").exists()


def test_low_memory_decoder_on_the_reference_known_answers(bce, tmp_path):
    """`bce -ds` needs no GPU: the reference's own archives decode to their inputs."""
    for i, v in enumerate(json.loads((GOLD / "kat.json").read_text())["vectors"]):
        arc = tmp_path / f"kat{i}.bce"
        arc.write_bytes(bytes.fromhex(v["archive_hex"]))
        code, out = run(bce, "-ds", tmp_path / f"out{i}", arc)
        data = eval(v["input_py"])
        assert code == 0 and f"Decompressed from {arc.stat().st_size} B -> {len(data)} B" in out
        assert (tmp_path / f"out{i}").read_bytes() == data


def test_bad_config_is_reported_and_ignored(bce, tmp_path):
    """bce.cpp:629-632: a config of the wrong size is not fatal.  Without a GPU the run then stops at the device
    (no CPU fallback); with one it goes on with the default table (checked in the GPU part)."""
    src = tmp_path / "in"
    src.write_bytes(b"hello world, hello world!")
    short = tmp_path / "short.bcc"
    short.write_bytes(b"\x01" * 100)
    code, out = run(bce, "-c", tmp_path / "a.bce", src, short)
    assert "Config not found or wrong size." in out
    code, out = run(bce, "-c", tmp_path / "a.bce", src, tmp_path / "missing.bcc")
    assert "Config not found or wrong size." in out
    wild = tmp_path / "wild.bcc"
    wild.write_bytes(b"\x09" * 288)
    code, out = run(bce, "-c", tmp_path / "a.bce", src, wild)
    assert "Config holds context bits above 5; ignored." in out


# ---- with a GPU ------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_compress_decompress_scan_end_to_end(bce, tmp_path):
    data = synth.generate("enwik-shaped", 300_000, 2).tobytes()
    src, arc, back = tmp_path / "in", tmp_path / "a.bce", tmp_path / "back"
    src.write_bytes(data)
    code, out = run(bce, "-c", arc, src)
    assert code == 0 and f"Compressed from {len(data)} B -> {arc.stat().st_size} B in" in out
    assert arc.read_bytes() == oracle.compress(data)
    for flag in ("-d", "-ds"):
        code, out = run(bce, flag, back, arc)
        assert code == 0 and f"Decompressed from {arc.stat().st_size} B -> {len(data)} B in" in out
        assert back.read_bytes() == data
        back.unlink()
    if oracle.have_ref():
        assert oracle.ref_decompress(arc.read_bytes()) == data

    g = json.loads((GOLD / "scan_markov2_200k.json").read_text())
    data = synth.generate("markov2-text", 200_000, 1).tobytes()
    src.write_bytes(data)
    cfg = tmp_path / "cfg.bcc"
    code, out = run(bce, "-s", cfg, src)
    assert code == 0 and f"Scanned {len(data)} B in" in out and out.count("Result size") == 9
    assert cfg.read_bytes().hex() == g["config_hex"]
    code, out = run(bce, "-c", arc, src, cfg)
    assert code == 0 and arc.stat().st_size == g["archive_bytes"]
    import hashlib
    assert hashlib.sha256(arc.read_bytes()).hexdigest() == g["archive_with_config_sha256"]
    code, out = run(bce, "-d", back, arc)
    assert code == 0 and back.read_bytes() == data


@pytest.mark.gpu
def test_bad_config_falls_back_to_the_default_table(bce, tmp_path):
    data = synth.generate("markov2-text", 50_000, 4).tobytes()
    src, arc = tmp_path / "in", tmp_path / "a.bce"
    src.write_bytes(data)
    code, out = run(bce, "-c", arc, src, tmp_path / "missing.bcc")
    assert code == 0 and "Config not found or wrong size." in out
    assert arc.read_bytes() == oracle.compress(data)


@pytest.mark.gpu
def test_exit_codes_of_device_failures_are_not_zero(bce, tmp_path):
    """A damaged archive that passes the size checks ends in a decoding error, never a crash."""
    data = synth.generate("markov2-text", 20_000, 5).tobytes()
    src, arc = tmp_path / "in", tmp_path / "a.bce"
    src.write_bytes(data)
    assert run(bce, "-c", arc, src)[0] == 0
    blob = bytearray(arc.read_bytes())
    blob[len(blob) // 2] ^= 0x55
    arc.write_bytes(bytes(blob))
    code, out = run(bce, "-d", tmp_path / "back", arc)
    assert code == 0 or "Decoding failed" in out          # the format has no checksum: garbage out or a clean error
