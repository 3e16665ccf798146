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
timeout 200 python -m pytest tests/test_gpu_ring.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -2
for k in 1 2 3 4 5 6; do
  echo "== real run $k"; BF_TIMING=1 $CLI --quiet --flow-out=/tmp/f.txt /tmp/stream.bin 2>&1 >/dev/null | grep -E "context|device objects|add_event loop|device ring, host|processing"
done
} > gpurun_out/cli_timing_r2w.txt 2>&1
cat gpurun_out/cli_timing_r2w.txt
