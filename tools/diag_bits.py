import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import better_flow_b200 as bf
from better_flow_b200 import synth
st = synth.make_stream(346, 260, 2.0e6, 0.05 * 8, seed=3)
sls = synth.cut_slices(st, 0.05)
c = bf.Context(260, 346, 3, max_events=len(st) + 64, max_slices=70, device=0)
s = sls[5]
ref = None
for G in (2, 4, 16, 32, 47, 48, 64, 148, 296):
    c.set_option("group_size", G)
    r = c.minimize(s.fr_x, s.fr_y, s.t_ns, scale=3, max_iter=10)
    if ref is None: ref = r
    d = [k for k in range(11) if r["model"][k] != ref["model"][k]]
    print("G", G, "used", c.get_option("group_size"), "iters", r["iters"], "diff fields", d, [(float(r["model"][k]).hex(), float(ref["model"][k]).hex()) for k in d[:2]])
