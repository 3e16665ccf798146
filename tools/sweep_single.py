"""Single-slice latency vs group size.  usage: sweep_single.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import better_flow_b200 as bf
from better_flow_b200 import synth
for name, cols, rows, rate, ss, mi in [("240C 10ms 30k", 240, 180, 3e6, 0.010, 10), ("240C 30ms 90k", 240, 180, 3e6, 0.030, -1),
                                      ("240C 50k (CLI window)", 240, 180, 5e6, 0.010, -1),
                                      ("346 50ms 100k", 346, 260, 2e6, 0.050, -1), ("640x480 200k", 640, 480, 10e6, 0.020, -1),
                                      ("1280x720 1M", 1280, 720, 100e6, 0.010, -1)]:
    st = synth.make_stream(cols, rows, rate, ss, seed=7)
    s = synth.cut_slices(st, ss)[0]
    ctx = bf.Context(rows, cols, 3, max_events=len(st) + 1024, max_slices=2, device=0)
    out = []
    for G, mg in ((0, 8), (16, 4), (16, 6), (16, 8), (16, 12), (16, 16), (8, 16), (24, 8), (32, 4), (32, 8)):
        ctx.set_option("group_size", max(G, 0))
        ctx.set_option("max_grow", mg)
        ctx.reset(); ctx.add(s.fr_x, s.fr_y, s.t_ns, 3, mi); ctx.run()
        ms = min(ctx.time_launches(5) / 5 for _ in range(3))
        r = ctx.result(0)
        out.append("G%d^%d %.3f" % (ctx.get_option("group_size"), mg, ms))
    print("%-24s n %7d iters %3d : ms/launch  %s" % (name, len(s.fr_x), r["iters"], "  ".join(out)), flush=True)
    ctx.close()
