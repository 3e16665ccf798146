"""Throughput of the OptimizerLocal path on a resident batch.  usage: prof_local.py slice_s n_slices [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import better_flow_b200 as bf
from better_flow_b200 import synth
ss, nsl = float(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
st = synth.make_stream(240, 180, 3e6, ss * nsl, seed=5)
sls = synth.cut_slices(st, ss)[:nsl]
ctx = bf.Context(180, 240, 3, max_events=len(st) + 1024, max_slices=len(sls) + 1, device=0)
for s in sls: ctx.add_local(s.fr_x, s.fr_y, s.t_ns, 3)
ctx.run()
ms = min(ctx.time_launches(2) / 2 for _ in range(reps))
res = [ctx.local_view(r) for r in ctx.results()]
nev = sum(r["n_events"] for r in res); steps = [r["steps"] for r in res]
print("local: slice %.3f n_slices %d G %d groups %d: %.3f ms/launch -> %.1f Mev/s, steps mean %.1f max %d, event-steps/s %.1f G" % (
    ss, len(sls), ctx.get_option("group_size"), ctx.get_option("n_groups"), ms, nev / ms / 1e3, np.mean(steps), max(steps),
    sum(r["n_events"] * r["steps"] for r in res) / ms / 1e6))
if os.environ.get("BF_CPU"):
    import time
    from oracle import ref, port
    t0 = time.perf_counter(); k = 0; ev = 0
    for s in sls[:6]:
        r = (ref.local_minimize if ref.available() else port.local_minimize)(s.fr_x, s.fr_y, s.t_ns, 3); k += 1; ev += len(s.fr_x)
    dt = time.perf_counter() - t0
    print("cpu reference OptimizerLocal: %d slices %.3f s -> %.2f Mev/s" % (k, dt, ev / dt / 1e6))
