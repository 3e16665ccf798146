#!/bin/bash
# round 2, visit u: ring tests (bf_ring_seed), then the default-mode CLI with the [timing] break-down, 5 runs each
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ring.py -m gpu -x -q 2>&1 | tail -3
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
for mode in "" "--stm-disable" "--no-device-ring"; do
  for k in 1 2 3 4 5; do
    echo "== mode '$mode' run $k"
    BF_TIMING=1 $CLI --quiet --flow-out=/tmp/flow_$k.txt $mode /tmp/stream.bin 2>&1 >/dev/null | grep -E "add_event loop|device ring, host|processing"
  done
done
echo "== null back end is not available here; host-only figure: see tests/cpu mock with BF_MOCK_NULL=1"
} > gpurun_out/cli_timing_r2u.txt 2>&1
grep -E "==|processing|host seconds" gpurun_out/cli_timing_r2u.txt | head -60
