"""Few-second GPU check of the three ways a slice enters a batch (bf_batch_add, _add_packed, _add_staged): same results,
and a coordinate outside the sensor is refused by all three."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import better_flow_b200 as bf
from better_flow_b200 import synth

st = synth.make_stream(240, 180, 3e6, 0.03, seed=77)
sls = synth.cut_slices(st, 0.01)[:3]
ctx = bf.Context(180, 240, 3, max_events=1 << 18, max_slices=8, device=0)
res = []
for how in ("add", "packed", "staged"):
    ctx.reset()
    off = 0
    for s in sls:
        ev = bf.pack_events(s.fr_x, s.fr_y, s.t_ns)
        if how == "add": ctx.add(s.fr_x, s.fr_y, s.t_ns, 3, 12)
        elif how == "packed": ctx.add_packed(ev, 3, 12)
        else:
            ctx.staging()[off:off + len(ev)] = ev
            ctx.add_staged(off, len(ev), 3, 12)
            off += len(ev)
    ctx.run()
    res.append([(r["rc"], r["iters"], r["model"].tobytes()) for r in ctx.results()])
assert res[0] == res[1] == res[2], "entry points disagree"
bad = bf.pack_events(sls[0].fr_x, sls[0].fr_y, sls[0].t_ns).copy()
bad["fr_x"][5] = 180
refused = 0
for how in ("packed", "staged"):
    ctx.reset()
    try:
        if how == "packed": ctx.add_packed(bad, 3, 12)
        else:
            ctx.staging()[:len(bad)] = bad
            ctx.add_staged(0, len(bad), 3, 12)
    except Exception as e:
        refused += "outside" in str(e)
assert refused == 2, refused
print("last check ok:", [r[1] for r in res[0]], "iterations; out-of-sensor batches refused")
ctx.close()
