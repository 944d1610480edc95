"""GPU: INTEGRATION.md, compiled and run.  oracle/_ref/bce_ref_gpu is the UNMODIFIED reference translation unit with
the binding INTEGRATION.md describes (oracle/ref_gpu_binding.cpp): GpuFrontEnd in place of RankFile, BCE::encode
fed from bce_gpu_cse_next into the reference's own AdaptiveCoder, unbwt::gpu as one more policy_unbwt under the
reference's own BCE::decode.  Its archives must equal `bce_ref -c` (the reference alone) and this repository's `bce`."""
import subprocess

import pytest

from bce_b200 import build, host, synth
from oracle import oracle

pytestmark = pytest.mark.gpu
BIN = oracle.HERE / "_ref" / "bce_ref_gpu"


def run(*args):
    r = subprocess.run([str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (args, r.stdout, r.stderr)
    return r.stdout


@pytest.mark.parametrize("kind,n,seed", [("enwik-shaped", 300_000, 2), ("mixed-binary", (1 << 20) + 5, 4), ("markov2-text", 7, 1)])
def test_reference_bound_to_the_gpu_library_writes_the_reference_archive(frontend, tmp_path, kind, n, seed):
    if not (BIN.exists() and oracle.have_ref()):
        pytest.skip("oracle/_ref not present")
    data = synth.generate(kind, n, seed).tobytes()
    src, arc, back = tmp_path / "in", tmp_path / "a.bce", tmp_path / "back"
    src.write_bytes(data)
    out = run(BIN, "-c", arc, src)
    assert f"Compressed from {n} B" in out
    blob = arc.read_bytes()
    assert blob == oracle.ref_compress(data), "reference + GPU front end != reference alone"
    assert blob == host.compress(frontend, data), "reference + GPU front end != this repository's bce -c"
    out = run(BIN, "-d", back, arc)                        # reference decoder, inverse BWT on the GPU (unbwt::gpu)
    assert back.read_bytes() == data
    back.unlink()
    run(build.BIN_BCE, "-d", back, arc)                    # and this repository's decoder on the same archive
    assert back.read_bytes() == data
