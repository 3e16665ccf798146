#!/bin/bash
# Round 2, final validation: every GPU test, smoke(), the bench line (both arms), the tool through the class surface.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_line.json 2> gpurun_out/bench.err
cat gpurun_out/bench_line.json; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_line_reference_arm.json 2> gpurun_out/bench_ref.err
cut -c1-300 gpurun_out/bench_line_reference_arm.json
BF_TIMING=1 timeout 100 ./better_flow_b200/bf_motion_compensator --quiet /tmp/stream.bin 2>&1 | grep timing | tail -4
