#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from better_flow_b200 import synth
st = synth.make_stream(240, 180, 3e6, 1.5, seed=1)
rec = np.zeros(len(st), dtype=np.dtype([("t", "<u8"), ("x", "<u2"), ("y", "<u2"), ("p", "<u4")]))
rec["t"], rec["x"], rec["y"], rec["p"] = st.t_ns, st.x, st.y, st.p
rec.tofile("/tmp/stream.bin")
PY
CLI=better_flow_b200/bf_motion_compensator
{
for mode in "" "--stm-disable"; do
for k in 1 2 3; do
  echo "== '$mode' run $k"; BF_TIMING=1 $CLI --quiet --flow-out=/tmp/f${k}_${mode:-default}.txt $mode /tmp/stream.bin 2>&1 >/dev/null | grep -E "device ring, host|processing"
done
done
BF_TIMING=1 $CLI --quiet --flow-out=/tmp/fh.txt --no-device-ring /tmp/stream.bin 2>&1 >/dev/null | grep -E "processing"
python - <<'PY'
import numpy as np
a = np.loadtxt("/tmp/f1_default.txt", ndmin=2); b = np.loadtxt("/tmp/fh.txt", ndmin=2)
rel = np.abs(a[:, 4:6] - b[:, 4:6]) / np.maximum(np.abs(b[:, 4:6]), 1e-300)
print("device ring vs host ring: %d slices, max rel (dx,dy) %.3g, iteration counts equal: %s" % (len(a), rel.max(), np.array_equal(a[:, 2], b[:, 2])))
PY
} > gpurun_out/cli_timing_r2za.txt 2>&1
cat gpurun_out/cli_timing_r2za.txt
