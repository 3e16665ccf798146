#!/bin/bash
# round 2, visit v: where the default-mode host time goes on the GPU box -- the tool with a NULL back end (the CPU
# test double, no compute, no CUDA) against the real one, plain and pinned to two cores
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
MOCK=tests/cpu/bf_motion_compensator_mock
{
grep -m1 "model name" /proc/cpuinfo; nproc; grep -m1 MHz /proc/cpuinfo
for k in 1 2 3 4 5; do
  echo "== null back end run $k"; BF_MOCK_NULL=1 BF_TIMING=1 $MOCK --quiet --flow-out=/tmp/f.txt /tmp/stream.bin 2>&1 >/dev/null | grep -E "processing"
done
for k in 1 2 3 4 5; do
  echo "== real run $k"; BF_TIMING=1 $CLI --quiet --flow-out=/tmp/f.txt /tmp/stream.bin 2>&1 >/dev/null | grep -E "device ring, host|processing"
done
for k in 1 2 3 4 5; do
  echo "== real, taskset 2 cores, run $k"; BF_TIMING=1 taskset -c 2,3 $CLI --quiet --flow-out=/tmp/f.txt /tmp/stream.bin 2>&1 >/dev/null | grep -E "device ring, host|processing"
done
for k in 1 2 3; do
  echo "== real, CUDA_DEVICE_MAX_CONNECTIONS=1 run $k"; CUDA_DEVICE_MAX_CONNECTIONS=1 BF_TIMING=1 $CLI --quiet --flow-out=/tmp/f.txt /tmp/stream.bin 2>&1 >/dev/null | grep -E "device ring, host|processing"
done
} > gpurun_out/cli_timing_r2v.txt 2>&1
cat gpurun_out/cli_timing_r2v.txt
