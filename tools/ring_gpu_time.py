"""How long the DEVICE needs for the default-mode chain of tools/cli_ring.py's stream (4.5 M events, a 50 k-event window
re-minimised every 20 k events, warm-started): every slice is enqueued through bf_ring_* as fast as Python can (pre-built
records, no result read-back before the end), then one sync.  If this is close to the tool's processing time, the tool
is bound by the device chain, not by its host loop."""
import os, sys, time
import ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import better_flow_b200 as bf
from better_flow_b200 import synth
from helpers import ring_slice
st = synth.make_stream(240, 180, 3e6, 1.5, seed=1)
fr_x, fr_y, ts = st.y, st.x, st.t_ns.astype(np.int64)
rec = np.zeros(len(ts), dtype=bf.RING_EVENT_DTYPE)
rec["fr_x"], rec["fr_y"], rec["timestamp"] = fr_x, fr_y, ts
consumed = list(range(20000, len(ts) + 1, 20000))
plan = []
for c in consumed:
    idx, start = ring_slice(ts, c)
    plan.append((c, len(idx), start))
ctx = bf.Context(180, 240, 3, max_events=50000 + 64, max_slices=64, device=0)
for chain in (1, 0):
    for rep in range(3):
        ring = bf.Ring(ctx, 50000, 256)
        lib, h = ring.lib, ring.h
        base = rec.ctypes.data
        ring.sync()
        t0 = time.perf_counter()
        fed = 0
        for c, n, start in plan:
            lib.bf_ring_push(h, C.c_void_p(base + 16 * fed), c - fed)
            fed = c
            lib.bf_ring_slice(h, n, C.c_uint64(start), 3, -1, chain)
        t1 = time.perf_counter()
        ring.sync()
        t2 = time.perf_counter()
        its = [ring.result(k)["iters"] for k in range(len(plan))]
        print("chain=%d: %d slices enqueued in %.1f ms, device done after %.1f ms (%.3f ms per slice, %.1f GD iterations per slice, %.1f Mev/s of stream)"
              % (chain, len(plan), 1e3 * (t1 - t0), 1e3 * (t2 - t0), 1e3 * (t2 - t0) / len(plan), np.mean(its), len(ts) / (t2 - t0) / 1e6))
        ring.close()
