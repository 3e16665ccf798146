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
# round 2: OptimizerLocal scale 5, the CLUSTER instance (batch + single slice), the compact upload, the device ring, debug images
ctx.reset()
ctx.add_local(sls[0].fr_x, sls[0].fr_y, sls[0].t_ns, 5)
ctx.add(sls[1].fr_x, sls[1].fr_y, sls[1].t_ns, 5, 3)
ctx.run()
print("local s5", [(r["rc"], r["iters"]) for r in ctx.results()])
for cs in (4, 16):
    ctx.set_option("cluster", cs)
    ctx.reset()
    for k, s in enumerate(sls):
        ctx.add(s.fr_x, s.fr_y, s.t_ns, (1, 3, 5)[k % 3], 3)
    ctx.run(True)
    print("cluster", cs, [r["iters"] for r in ctx.results()], ctx.get_option("group_size"))
ctx.set_option("cluster", 0)
for rep in range(3):
    ctx.reset()
    for k, s in enumerate(sls):
        ctx.add_delta(bf.pack_events(s.fr_x[:len(s.fr_x) - k], s.fr_y[:len(s.fr_y) - k], s.t_ns[:len(s.t_ns) - k]), 3, 3)
    ctx.set_option("upload_chunks", 1 + 2 * rep)
    ctx.run_streamed(False); ctx.sync()
print("delta", [r["iters"] for r in ctx.results()])
for rc in (0, 8):
    ctx.set_option("ring_cluster", rc)
    ring = bf.Ring(ctx, 6000, 4)
    tk = []
    for k in range(6):
        lo, hi = 2000 * k, 2000 * (k + 1)
        ring.push(st.y[lo:hi], st.x[lo:hi], st.t_ns[lo:hi])
        tk.append(ring.slice(min(hi, 5999), max(0, int(st.t_ns[hi - 1]) - 3_000_000), 3, 3, True))
    print("ring", rc, [ring.result(t)["iters"] for t in tk[-4:]])
    ring.close()
ctx.set_option("ring_cluster", 0)
s = sls[0]
pi, avg = ctx.projection_img(s.fr_x.astype(float) - 0.7, s.fr_y.astype(float) + 0.4, 5)
ci = ctx.color_time_img(s.fr_x.astype(float) - 0.7, s.fr_y.astype(float) + 0.4, s.t_ns, 5)
print("images", int(pi.sum()), avg, int(ci.sum()))
s = sls[0]
su_w, su_h = 3 * 179, 3 * 239
img = ctx.time_img(s.fr_x.astype(float), s.fr_y.astype(float), s.t_ns, su_w, su_h, 3, 2, 2)
m = ctx.fast_model(s.fr_x.astype(float), s.fr_y.astype(float), s.t_ns, su_w, su_h, 3, 2, 2)
m2 = ctx.model_from_image(img)
print(img.sum(), m[6], m2[6])
ctx.close()
