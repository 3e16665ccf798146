#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "delta or streamed" ) 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/bench_line.json 2> gpurun_out/bench.err
cat gpurun_out/bench_line.json; tail -3 gpurun_out/bench.err
