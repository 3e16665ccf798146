"""The ring entry points added late in round 2 under compute-sanitizer: bf_ring_reserve / bf_ring_commit (partial commits,
a commit larger than the ring, staging wrap), bf_ring_seed, chained and unchained slices."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import better_flow_b200 as bf
from better_flow_b200 import synth
st = synth.make_stream(240, 180, 1.5e6, 0.04, seed=5)
fr_x, fr_y, ts = st.y, st.x, st.t_ns.astype(np.int64)
ctx = bf.Context(180, 240, 3, max_events=9000, max_slices=8, device=0)
ring = bf.Ring(ctx, 6000, 4)
ring.seed(np.array([90.0, 120.0, 0.01, -0.02, 0.0, 0.0, 100.0, 0.3, -0.2, 0.001, 0.0005]))
tk, fed = [], 0
for k in range(24):                       # 24 x 2500 events through a 6000-event ring; the 262144-entry staging buffer is
    lo, hi = fed, fed + 2500              # reserved 32768 at a time, so it wraps after 8 reservations
    ring.push_in_place(fr_x[lo:lo + 900], fr_y[lo:lo + 900], ts[lo:lo + 900], reserve=32768)
    ring.push_in_place(fr_x[lo + 900:hi], fr_y[lo + 900:hi], ts[lo + 900:hi], reserve=32768)
    fed = hi
    tk.append(ring.slice(min(fed, 5999), max(0, int(ts[fed - 1]) - 3_000_000), 3, 3, k % 3 != 0))
print("reserve/commit", [ring.result(t)["iters"] for t in tk[-4:]], ring.pushed)
ring.close()
ring = bf.Ring(ctx, 4000, 4)
ring.push_in_place(fr_x[:30000], fr_y[:30000], ts[:30000])          # one commit, 7.5 x the ring
print("oversized commit", ring.pushed, ring.result(ring.slice(4000, int(ts[26000]), 3, 3, False))["iters"])
ring.close()
ctx.close()
