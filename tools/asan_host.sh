#!/bin/bash
# The host side (DVS_flow mirror, tool, event I/O) under AddressSanitizer + UBSan, linked against the CPU test double of
# the C ABI (tests/cpu/mock_bf_cuda.cpp: its rings are freed with their context, like the library's, so a stale ring or
# staging pointer is a use-after-free the sanitizer sees).  No GPU.  usage: bash tools/asan_host.sh
set -e
cd "$(dirname "$0")/.."
make -s -C oracle port
FLAGS="-std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off -pthread -Ibetter_flow_b200/include -Iinclude"
LINK="-Loracle -lbf_oracle -Wl,-rpath,$PWD/oracle"
g++ $FLAGS better_flow_b200/src/bf_motion_compensator.cpp tests/cpu/mock_bf_cuda.cpp $LINK -o /tmp/bf_cli_asan
g++ $FLAGS tests/cpu/asan_stream_main.cpp tests/cpu/stream_shim.cpp tests/cpu/mock_bf_cuda.cpp $LINK -o /tmp/bf_stream_asan
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from better_flow_b200 import synth
st = synth.make_stream(240, 180, 1.0e6, 0.25, seed=3)
rec = np.zeros(len(st), dtype=np.dtype([("t", "<u8"), ("x", "<u2"), ("y", "<u2"), ("p", "<u4")]))
rec["t"], rec["x"], rec["y"], rec["p"] = st.t_ns, st.x, st.y, st.p
rec.tofile("/tmp/bf_asan.bin"); st.to_text("/tmp/bf_asan.txt")
PY
export ASAN_OPTIONS=detect_leaks=1
fail=0
while read -r args; do
  if /tmp/bf_cli_asan $args > /tmp/bf_asan_out.txt 2> /tmp/bf_asan_err.txt; then echo "ok    tool $args"; else echo "FAIL  tool $args"; grep -E "ERROR|runtime error|SUMMARY" /tmp/bf_asan_err.txt | head -5; fail=1; fi
done <<'ARGS'
--quiet --max-iter=3 /tmp/bf_asan.bin
--quiet --max-iter=3 --stm-disable /tmp/bf_asan.bin
--quiet --max-iter=3 --no-device-ring /tmp/bf_asan.bin
--quiet --max-iter=3 -o /tmp/bf_asan_uv.txt /tmp/bf_asan.bin
--quiet --max-iter=3 --stm-disable --batch=4 -o /tmp/bf_asan_uv2.txt /tmp/bf_asan.txt
--quiet --max-iter=2 --refresh-event-count=40000 --refresh-time=1 /tmp/bf_asan.bin
--quiet --max-iter=2 --max-events=20000 --refresh-event-count=45000 --refresh-time=1 /tmp/bf_asan.bin
--max-iter=2 /tmp/bf_asan.txt
ARGS
if /tmp/bf_stream_asan > /tmp/bf_asan_out.txt 2> /tmp/bf_asan_err.txt; then echo "ok    DVS_flow: host ring, device ring, mode switched mid-stream, context re-created mid-stream (x stm on/off x 2 buffer configurations)"; else echo "FAIL  stream scenarios"; grep -E "ERROR|runtime error|SUMMARY" /tmp/bf_asan_err.txt | head -5; fail=1; fi
exit $fail
