"""Small workload for compute-sanitizer: a few slices, few iterations, all scales, writeout, stage API."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import better_flow_b200 as bf
from better_flow_b200 import synth
st = synth.make_stream(240, 180, 1.5e6, 0.03, seed=3)
sls = synth.cut_slices(st, 0.005)
ctx = bf.Context(180, 240, 5, max_events=len(st) + 64, max_slices=16, device=0)
for G in (0, 8):
    ctx.set_option("group_size", G)
    ctx.reset()
    for k, s in enumerate(sls):
        ctx.add(s.fr_x, s.fr_y, s.t_ns, (1, 3, 5)[k % 3], 3)
    ctx.run(True)
    print([r["iters"] for r in ctx.results()], ctx.get_option("group_size"))
    ctx.run_streamed(False); ctx.sync()
# OptimizerLocal slices mixed with rolling ones, helpers joining from the first iterations (G = 2: 148 groups, 6 slices)
ctx.set_option("group_size", 2)
ctx.reset()
for k, s in enumerate(sls):
    if k % 2: ctx.add_local(s.fr_x, s.fr_y, s.t_ns, (1, 3)[k % 4 == 1])
    else: ctx.add(s.fr_x, s.fr_y, s.t_ns, (1, 3, 5)[k % 3], 4)
ctx.run(True)
print([(r["rc"], r["iters"]) for r in ctx.results()])
ctx.set_option("group_size", 0)
s = sls[0]
su_w, su_h = 3 * 179, 3 * 239
img = ctx.time_img(s.fr_x.astype(float), s.fr_y.astype(float), s.t_ns, su_w, su_h, 3, 2, 2)
m = ctx.fast_model(s.fr_x.astype(float), s.fr_y.astype(float), s.t_ns, su_w, su_h, 3, 2, 2)
m2 = ctx.model_from_image(img)
print(img.sum(), m[6], m2[6])
ctx.close()
