"""Run one resident batch a few times (for ncu / quick timing).  Usage: prof_batch.py slice_s max_iter n_slices G reps [rate]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import better_flow_b200 as bf
from better_flow_b200 import synth

ss, mi, nsl, G, reps = float(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
rate = float(sys.argv[6]) if len(sys.argv) > 6 else 3e6
cols, rows = (int(sys.argv[7]), int(sys.argv[8])) if len(sys.argv) > 8 else (240, 180)
st = synth.make_stream(cols, rows, rate, ss * nsl, seed=5)
sls = synth.cut_slices(st, ss)
ctx = bf.Context(rows, cols, 3, max_events=len(st) + 1024, max_slices=len(sls) + 1, device=0)
ctx.set_option("ctas_per_sm", int(os.environ.get("BF_CPS", "2")))
ctx.set_option("group_size", G)
if os.environ.get("BF_TAILHELP") is not None:
    try: ctx.set_option("tail_help", int(os.environ["BF_TAILHELP"]))
    except Exception: pass
for s in sls: ctx.add(s.fr_x, s.fr_y, s.t_ns, 3, mi)
ctx.run()
ms = ctx.time_launches(reps) / reps
res = ctx.results()
nev = sum(r["n_events"] for r in res); its = [r["iters"] for r in res]
P = res[0]["img_rows"] * res[0]["img_cols"]
alg = sum(r["iters"] * (40 * r["n_events"] + 16 * P) for r in res)
print("slice %.3f mi %d n_slices %d G %d groups %d: %.3f ms/launch -> %.1f Mev/s, iters mean %.1f max %d, alg GB/s %.1f" % (
    ss, mi, len(sls), ctx.get_option("group_size"), ctx.get_option("n_groups"), ms, nev / ms / 1e3, np.mean(its), max(its), alg / ms / 1e6))

if os.environ.get("BF_DUMP_ITERS"):
    import json
    json.dump({"iters": its, "n": [r["n_events"] for r in res]}, open(os.environ["BF_DUMP_ITERS"], "w"))

if os.environ.get("BF_PROFILE"):
    ctx.set_option("profile", 1)
    ctx.launch(); ctx.sync()
    pf = ctx.debug_profile().astype(np.float64)
    names = ["event", "barA", "sums", "cells", "reduce", "barB", "serial", "prologue", "final", "iters", "slices", "barA_spin", "barB_spin", "total", "sums_loads"]
    tot = pf[:, 13].mean()
    print("per-CTA mean cycles (%% of total %.0f):" % tot)
    for k, nm in enumerate(names):
        if nm in ("iters", "slices"): print("  %-10s %10.1f" % (nm, pf[:, k].mean())); continue
        print("  %-10s %12.0f  %5.1f%%   per-iter %8.0f cyc" % (nm, pf[:, k].mean(), 100 * pf[:, k].mean() / tot, pf[:, k].mean() / max(pf[:, 9].mean(), 1)))
    print("  CTA totals min/max cycles: %.0f / %.0f" % (pf[:, 13].min(), pf[:, 13].max()))
    if os.environ.get("BF_DUMP_ITERS"):
        np.save(os.environ["BF_DUMP_ITERS"] + ".pf.npy", pf)
