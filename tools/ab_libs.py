"""Same-box A/B of several builds of libbf_cuda.so on the bench workload (interleaved subprocess runs).
usage: ab_libs.py n_slices reps lib1.so lib2.so ...   (paths relative to the repo root)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
nsl, reps, libs = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3:]
CODE = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np
import better_flow_b200 as bf
from better_flow_b200 import synth
nsl = %d
st = synth.make_stream(240, 180, 3e6, 0.03 * nsl, seed=100)
sls = synth.cut_slices(st, 0.03)[:nsl]
ctx = bf.Context(180, 240, 3, max_events=len(st) + 1024, max_slices=len(sls) + 1, device=0)
for kv in %r.split(","):
    if kv: ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
for s in sls: ctx.add(s.fr_x, s.fr_y, s.t_ns, 3, -1)
ctx.run()
ts = [ctx.time_launches(2) / 2 for _ in range(4)]
res = ctx.results()
nev = sum(r["n_events"] for r in res)
chk = sum(float(r["model"][7]) + float(r["model"][8]) for r in res)
print("%%.3f %%.3f %%.1f %%d %%.17g" %% (min(ts), float(np.median(ts)), nev / min(ts) / 1e3, sum(r["iters"] for r in res), chk))
'''
out = {l: [] for l in libs}
for rep in range(reps):
    for l in libs:
        l, _, OPTS = l.partition(":")      # lib.so:key=value,key=value
        env = dict(os.environ, BF_LIB_PATH=os.path.join(ROOT, l))
        code = CODE % (ROOT, nsl, OPTS)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        out[l + ":" + OPTS if OPTS else l].append(r.stdout.strip() or r.stderr.strip()[-300:])
for l in libs:
    print(l)
    for o in out[l]:
        print("   min_ms median_ms Mev/s iters checksum:", o)
