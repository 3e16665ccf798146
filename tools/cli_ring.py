"""Default-mode throughput of the drop-in CLI through the DVS_flow class surface: the events are
read up front (binary input = the tool's --bufferize-file path), then the add_event loop + every slice + every model
read-back is timed (steady_clock, the tool's own "[timing] processing" line under BF_TIMING).  Device ring (default)
against the host ring (--no-device-ring), warm-start chain and --stm-disable; the reference's own tool on the same
stream as text when oracle/_ref travelled."""
import json, os, re, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from better_flow_b200 import synth
CLI = os.path.join(ROOT, "better_flow_b200", "bf_motion_compensator")
REF = os.path.join(ROOT, "oracle", "_ref", "bf_motion_compensator_ref")
dur = float(sys.argv[1]) if len(sys.argv) > 1 else 1.5
with_ref = len(sys.argv) > 2 and sys.argv[2] == "ref"
st = synth.make_stream(240, 180, 3e6, dur, seed=1)
rec = np.zeros(len(st), dtype=np.dtype([("t", "<u8"), ("x", "<u2"), ("y", "<u2"), ("p", "<u4")]))
rec["t"], rec["x"], rec["y"], rec["p"] = st.t_ns, st.x, st.y, st.p
binf = "/tmp/stream.bin"
rec.tofile(binf)
print("stream: %d events, %.2f s of sensor time" % (len(st), dur))
out = {"events": len(st)}
PROC = re.compile(r"\[timing\] processing (\d+) events in ([0-9.e+-]+) s = ([0-9.e+-]+) Mev/s, slices (\d+)")
def run(extra, label, key, env=None):
    best = None
    for rep in range(3):
        flow = "/tmp/flow_%s.txt" % key
        r = subprocess.run([CLI, "--quiet", "--flow-out=" + flow] + extra + [binf], capture_output=True, text=True, env=dict(os.environ, BF_TIMING="1", **(env or {})))
        m = PROC.search(r.stderr)
        if r.returncode != 0 or not m:
            print(label, "FAILED", r.stderr[-400:]); return None
        h = re.search(r"add_event loop returned after ([0-9.e+-]+) s", r.stderr)
        v = (float(m.group(2)), float(m.group(3)), int(m.group(4)), float(h.group(1)) if h else -1.0)
        best = v if best is None or v[0] < best[0] else best
    print("%-52s processing %.4f s  %8.1f Mev/s  %d slices  (add_event loop returned after %.4f s)" % (label, best[0], best[1], best[2], best[3]))
    out[key] = {"seconds": best[0], "mevs": best[1], "slices": best[2], "host_loop_seconds": best[3]}
    return np.loadtxt(flow, ndmin=2)
a = run([], "default mode, device ring", "default_ring")
for cs in (8, 16):
    run([], "default mode, device ring, cluster of %d" % cs, "default_ring_cluster%d" % cs, env={"BF_RING_CLUSTER": str(cs)})
    run(["--stm-disable"], "--stm-disable, device ring, cluster of %d" % cs, "stm_ring_cluster%d" % cs, env={"BF_RING_CLUSTER": str(cs)})
b = run(["--no-device-ring"], "default mode, host ring (round 1 path)", "default_host")
c = run(["--stm-disable"], "--stm-disable, device ring", "stm_ring")
d = run(["--stm-disable", "--no-device-ring"], "--stm-disable, host ring", "stm_host")
e = run(["--stm-disable", "--batch=16"], "--stm-disable --batch=16", "stm_batch16")
if a is not None and b is not None:
    rel = np.abs(a[:, 4:6] - b[:, 4:6]) / np.maximum(np.abs(b[:, 4:6]), 1e-300)
    out["ring_vs_host_max_rel"] = float(rel.max()); out["ring_vs_host_iters_equal"] = bool(np.array_equal(a[:, 2], b[:, 2]))
    print("device ring vs host ring: %d slices, max rel (dx,dy) %.3g, iteration counts equal: %s" % (len(a), rel.max(), out["ring_vs_host_iters_equal"]))
if c is not None and d is not None:
    rel = np.abs(c[:, 4:6] - d[:, 4:6]) / np.maximum(np.abs(d[:, 4:6]), 1e-300)
    print("stm-disable ring vs host: max rel %.3g" % rel.max())
if with_ref and os.path.exists(REF):
    txt = "/tmp/stream.txt"
    st.to_text(txt)
    t0 = time.perf_counter()
    r = subprocess.run([REF, "--bufferize-file", txt], capture_output=True, text=True)
    m = re.search(r"Toatal flow elapsed: ([0-9.e+-]+) sec", r.stdout)
    print("reference tool, default mode, --bufferize-file: total wall %.2f s, its own std::clock figure %s s" % (time.perf_counter() - t0, m.group(1) if m else "?"))
    out["reference_cpu_clock_s"] = float(m.group(1)) if m else None
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "cli_ring.json"), "w"))
