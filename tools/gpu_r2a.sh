#!/bin/bash
# Round 2, visit A: GPU tests (new parity tests included), bench line, A/B of the event-load variants.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q -s ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_line.json 2> gpurun_out/bench.err
cat gpurun_out/bench_line.json; tail -3 gpurun_out/bench.err
timeout 600 python tools/ab_libs.py 592 3 better_flow_b200/libbf_cuda.so build/libbf_evld0.so build/libbf_evld2.so > gpurun_out/ab.txt 2>&1
cat gpurun_out/ab.txt
