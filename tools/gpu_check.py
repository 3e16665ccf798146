"""Verbose first-contact check for the GPU box (prints instead of asserting)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import better_flow_b200 as bf
from better_flow_b200 import synth
from oracle import port

def rel(a, b): return np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(b), 1e-300)

ctx = bf.Context(180, 240, 5, max_events=1 << 22, max_slices=256, device=0)
print("sms", ctx.get_option("sms"), "smem", ctx.get_option("smem_bytes"), "G", ctx.get_option("group_size"),
      "groups", ctx.get_option("n_groups"), "img MB", ctx.get_option("image_bytes") / 1e6)
st = synth.make_stream(240, 180, 3e6, 0.06, seed=1)
sl = synth.cut_slices(st, 0.01)[0]
n = len(sl.fr_x)
rng = np.random.default_rng(0)
# project
pr_x = sl.fr_x + rng.normal(0, 2, n); pr_y = sl.fr_y + rng.normal(0, 2, n)
args = (-0.043, 0.081, 91.3, 118.7, 3.1e-5, -2.2e-4)
g = ctx.project(sl.fr_x, sl.fr_y, sl.t_ns, pr_x, pr_y, *args)
w = port.project(sl.fr_x, sl.fr_y, sl.t_ns, pr_x, pr_y, *args)
print("project: bit-exact", [bool(np.array_equal(a, b)) for a, b in zip(g, w)], "maxdiff", [float(np.max(np.abs(a - b))) for a, b in zip(g, w)])
# images
for scale in (3, 1, 5):
    su = port.setup_slice(sl.fr_x, sl.fr_y, 180, 240, scale)
    a = (pr_x, pr_y, sl.t_ns, su.wsize_x, su.wsize_y, scale, int(su.x_shift), int(su.y_shift))
    img = ctx.time_img(*a)
    ex = port.time_img(*a, accum_mode=1); rf = port.time_img(*a, accum_mode=0)
    print("scale", scale, "time_img exact-equal", bool(np.array_equal(img, ex)), "ndiff", int((img != ex).sum()), "max|d| vs ref", float(np.max(np.abs(img - rf))), "nnz", int((img > 0).sum()), int((ex > 0).sum()))
    o7, gx, gy = ctx.fast_model(*a, want_grad=True)
    w7, wgx, wgy = port.model(ex, want_grad=True)
    print("   model", o7, "\n   want ", w7, "\n   grad equal", bool(np.array_equal(gx, wgx)), bool(np.array_equal(gy, wgy)), "ndiff", int((gx != wgx).sum()))
# minimise
for (ss, mi, sc) in [(0.01, 10, 3), (0.01, -1, 3), (0.03, -1, 3), (0.01, 25, 1)]:
    s = synth.cut_slices(st, ss)[0]
    t0 = time.time(); got = ctx.minimize(s.fr_x, s.fr_y, s.t_ns, scale=sc, max_iter=mi, want_events=True); dt = time.time() - t0
    ex = port.minimize(s.fr_x, s.fr_y, s.t_ns, scale=sc, max_iter=mi, accum_mode=1, want_events=True)
    rf = port.minimize(s.fr_x, s.fr_y, s.t_ns, scale=sc, max_iter=mi, accum_mode=0)
    print("minimize", ss, mi, sc, "n", len(s.fr_x), "rc", got["rc"], "iters", got["iters"], ex["iters"], rf["iters"], "%.1f ms" % (dt * 1e3))
    print("   got  ", got["model"]); print("   exact", ex["model"])
    print("   rel vs exact", rel(got["model"][7:11], ex["model"][7:11]), "rel vs ref", rel(got["model"][7:11], rf["model"][7:11]))
    print("   pr maxdiff", float(np.max(np.abs(got["pr_x"] - ex["pr_x"]))), "nx maxdiff", float(np.max(np.abs(got["nx"] - ex["nx"]))), "div", got["dividers"], ex["dividers"])
# batch timing
for (ss, mi, nsl) in [(0.01, 10, 64), (0.03, -1, 64)]:
    stb = synth.make_stream(240, 180, 3e6, ss * nsl, seed=5)
    sls = synth.cut_slices(stb, ss)
    for G in (0, 4, 8, 16, 37):
        ctx.set_option("group_size", G)
        ctx.reset()
        for s in sls: ctx.add(s.fr_x, s.fr_y, s.t_ns, 3, mi)
        ctx.run()
        ms = ctx.time_launches(3) / 3
        res = ctx.results()
        nev = sum(r["n_events"] for r in res); its = [r["iters"] for r in res]
        print("batch slice %.3f mi %d n_slices %d G %d groups %d: %.3f ms/launch -> %.1f Mev/s, iters mean %.1f max %d, ev-iters/s %.2f G" % (
            ss, mi, len(sls), ctx.get_option("group_size"), ctx.get_option("n_groups"), ms, nev / ms / 1e3, np.mean(its), max(its),
            sum(r["n_events"] * r["iters"] for r in res) / ms / 1e6))
ctx.close()
print("done")
