"""Resident-batch timing of every BASELINE.json configuration (parity for these is in tests/test_gpu_baseline_sizes.py).
Prints one line per config: Mev/s, ms per launch, iterations, launch geometry."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import better_flow_b200 as bf
from better_flow_b200 import synth

CONFIGS = [
    # name, cols, rows, rate, slice_s, n_slices, max_iter
    ("cfg1 DAVIS-240C 10 ms x1, 10 GD iters", 240, 180, 3e6, 0.010, 1, 10),
    ("cfg2 DAVIS-240C 30 ms x592, to convergence", 240, 180, 3e6, 0.030, 592, -1),
    ("cfg3 DAVIS-346 50 ms x64, to convergence", 346, 260, 2e6, 0.050, 64, -1),
    ("cfg4 640x480 20 ms (200k ev) x32, to convergence", 640, 480, 10e6, 0.020, 32, -1),
    ("cfg4' 640x480 20 ms x1", 640, 480, 10e6, 0.020, 1, -1),
    ("cfg5 1280x720 10 ms (1M ev) x8, to convergence", 1280, 720, 100e6, 0.010, 8, -1),
    ("cfg5' 1280x720 10 ms x1", 1280, 720, 100e6, 0.010, 1, -1),
]
for name, cols, rows, rate, ss, nsl, mi in CONFIGS:
    st = synth.make_stream(cols, rows, rate, ss * nsl, seed=7)
    sls = synth.cut_slices(st, ss)[:nsl]
    ctx = bf.Context(rows, cols, 3, max_events=len(st) + 1024, max_slices=len(sls) + 1, device=0)
    for s in sls: ctx.add(s.fr_x, s.fr_y, s.t_ns, 3, mi)
    ctx.run()
    ms = min(ctx.time_launches(3) / 3 for _ in range(3))
    res = ctx.results()
    nev = sum(r["n_events"] for r in res); its = [r["iters"] for r in res]
    P = res[0]["img_rows"] * res[0]["img_cols"]
    alg = sum(r["iters"] * (40 * r["n_events"] + 16 * r["img_rows"] * r["img_cols"]) for r in res)
    print("%-52s %9.1f Mev/s  %8.3f ms/launch  iters mean %5.1f max %3d  G %3d x %3d groups  rc0 %d/%d  alg %.0f GB/s" % (
        name, nev / ms / 1e3, ms, np.mean(its), max(its), ctx.get_option("group_size"), ctx.get_option("n_groups"),
        sum(r["rc"] == 0 for r in res), len(res), alg / ms / 1e6), flush=True)
    ctx.close()
