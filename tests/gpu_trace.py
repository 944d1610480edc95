"""Developer aid: one warm-up pass, then one resident front-end pass with BCE_GPU_TRACE=1
(launch-by-launch timing on stderr) and the stage timings."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bce_b200 import Frontend, synth  # noqa: E402

kind, n, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
d = synth.generate(kind, n, seed)
fe = Frontend(0)
if os.environ.get("BCE_TRACE_EMIT") == "coder":
    from bce_b200.gpu import EMIT_CODER
    fe.set_emit_mode(EMIT_CODER)
fe.stage_input(d)
if os.environ.get("BCE_TRACE_WARM", "1") != "0":
    fe.front_resident()
os.environ["BCE_GPU_TRACE"] = "1"
fe.front_resident()
os.environ["BCE_GPU_TRACE"] = "0"
st = fe.stats()
print(json.dumps({k: v for k, v in st.items() if k.startswith("ms_") or k.startswith("cse_") or k.startswith("sort_")}))
