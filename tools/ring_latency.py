"""Where one slice of the default-mode chain spends its ~0.27 ms on the device: in-kernel clock64 phase counters of the
last slice of a warm chain (profile option), for a few launch shapes.  Under ncu (--metrics gpu__time_duration.sum) the
same script gives the per-kernel durations (ring build kernel, memset, persistent kernel).
usage: ring_latency.py [n_slices]"""
import os, sys, time
import ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import better_flow_b200 as bf
from better_flow_b200 import synth
from helpers import ring_slice
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 40
st = synth.make_stream(240, 180, 3e6, 0.02 + NS * 0.0067, seed=1)
fr_x, fr_y, ts = st.y, st.x, st.t_ns.astype(np.int64)
rec = np.zeros(len(ts), dtype=bf.RING_EVENT_DTYPE)
rec["fr_x"], rec["fr_y"], rec["timestamp"] = fr_x, fr_y, ts
plan = []
for c in list(range(20000, len(ts) + 1, 20000))[:NS]:
    idx, start = ring_slice(ts, c)
    plan.append((c, len(idx), start))
ctx = bf.Context(180, 240, 3, max_events=50000 + 64, max_slices=64, device=0)
names = ["event", "bar_a", "scan", "cells", "reduce", "bar_b", "serial", "prologue", "final", "iters", "slices", "spin_a", "spin_b", "total"]
def chain(label, **opts):
    for k, v in opts.items():
        ctx.set_option(k, v)
    best = None
    for rep in range(3):
        ring = bf.Ring(ctx, 50000, 256)
        lib, h, base = ring.lib, ring.h, rec.ctypes.data
        ring.sync()
        t0 = time.perf_counter()
        fed = 0
        for c, n, start in plan:
            lib.bf_ring_push(h, C.c_void_p(base + 16 * fed), c - fed)
            fed = c
            lib.bf_ring_slice(h, n, C.c_uint64(start), 3, -1, 1)
        ring.sync()
        dt = time.perf_counter() - t0
        its = np.mean([ring.result(k)["iters"] for k in range(len(plan))])
        best = dt if best is None else min(best, dt)
        ring.close()
    print("%-34s %.3f ms per slice (%d slices, %.1f GD iterations per slice)" % (label, 1e3 * best / len(plan), len(plan), its), flush=True)
chain("default launch")
if os.environ.get("BF_PROFILE_PHASES"):
    ctx.set_option("profile", 1)
    chain("default launch, phase counters on")
    p = ctx.debug_profile()
    act = p[p[:, 13] > 0]
    lead = act[np.argmax(act[:, 7])]          # the CTA with the longest prologue = a member of the group that owned the slice
    clk = 1.0 / 1.9e3                         # us per cycle at ~1.9 GHz
    print("last slice: %d CTAs took part; leader-side CTA cycles -> us:" % len(act))
    print("  " + "  ".join("%s %.1f" % (names[k], lead[k] * clk) for k in (7, 0, 1, 2, 3, 4, 5, 6, 8, 13)) + "  iters %d" % lead[9])
    ctx.set_option("profile", 0)
if os.environ.get("BF_WIDE_GROUPS"):
    for G, mg, th in ((32, 2, 1), (32, 1, 0), (48, 1, 0), (64, 1, 0), (96, 1, 0), (24, 3, 1), (16, 4, 1)):
        chain("group_size %d, max_grow %d, tail_help %d" % (G, mg, th), group_size=G, max_grow=mg, tail_help=th)
    sys.exit(0)
for G, mg in ((4, 8), (4, 16), (8, 8), (16, 4), (16, 8), (2, 32)):
    chain("group_size %d, max_grow %d" % (G, mg), group_size=G, max_grow=mg)
ctx.set_option("group_size", 0); ctx.set_option("max_grow", 8)
chain("tail_help off (one group of 4)", tail_help=0)
ctx.set_option("tail_help", 1)
for cs in (8, 16):
    chain("ring_cluster %d" % cs, ring_cluster=cs)
