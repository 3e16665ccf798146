"""End-to-end wall time of the drop-in CLI against the reference's own CLI (oracle/_ref) on one text stream."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from better_flow_b200 import synth
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "better_flow_b200", "bf_motion_compensator")
REF = os.path.join(ROOT, "oracle", "_ref", "bf_motion_compensator_ref")
dur = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
st = synth.make_stream(240, 180, 3e6, dur, seed=1)
txt = "/tmp/stream.txt"
st.to_text(txt)
print("stream: %d events, %.2f s" % (len(st), dur))
def run(cmd, label):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, BF_TIMING="1"))
    dt = time.perf_counter() - t0
    tail = [l for l in (r.stderr + r.stdout).splitlines() if "slices" in l or "elapsed" in l.lower()][-2:]
    print("%-44s %.3f s  rc %d  %s" % (label, dt, r.returncode, " | ".join(tail)[:160]))
    for l in r.stderr.splitlines():
        if l.startswith("[timing]"): print("      " + l)
for extra, lab in ([], "default (warm start, 50k/200ms window)"), (["--stm-disable"], "--stm-disable"), (["--stm-disable", "--batch=16"], "--stm-disable --batch=16"):
    run([CLI, "--quiet"] + extra + [txt], "ours " + lab)
if os.path.exists(REF):
    run([REF, txt], "reference default")
    run([REF, "--stm-disable", txt], "reference --stm-disable")
