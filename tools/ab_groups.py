"""Same-box A/B of group sizes / CTAs per SM on one resident batch (interleaved repeats)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import better_flow_b200 as bf
from better_flow_b200 import synth
ss, mi, nsl = float(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
variants = [tuple(int(v) for v in a.split(":")) for a in sys.argv[4].split(",")]   # cps:G
st = synth.make_stream(240, 180, 3e6, ss * nsl, seed=5)
sls = synth.cut_slices(st, ss)
ctxs = {}
for cps, G in variants:
    c = bf.Context(180, 240, 3, max_events=len(st) + 1024, max_slices=len(sls) + 1, device=0)
    c.set_option("ctas_per_sm", cps); c.set_option("group_size", G)
    for s in sls: c.add(s.fr_x, s.fr_y, s.t_ns, 3, mi)
    c.run()
    ctxs[(cps, G)] = c
nev = sum(len(s.fr_x) for s in sls)
times = {k: [] for k in ctxs}
for rep in range(5):
    for k, c in ctxs.items():
        times[k].append(c.time_launches(2) / 2)
for k, v in times.items():
    print("cps %d G %3d groups %3d: min %.3f ms median %.3f ms -> %.1f Mev/s (best)" % (
        k[0], k[1], ctxs[k].get_option("n_groups"), min(v), float(np.median(v)), nev / min(v) / 1e3))
