#!/bin/bash
# round 2, last visit: the whole GPU suite, smoke, sanitizer over the late ring entry points, one bench line per arm
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash tools/gpu_r2za.sh | tail -12
timeout 300 python bench.py > gpurun_out/bench_line.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench_line.json
