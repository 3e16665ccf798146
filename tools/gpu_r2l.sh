#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_ring.py -m gpu -x -q ) 2>&1 | tail -3
timeout 600 python tools/ab_libs.py 592 3 build/libbf_prev.so better_flow_b200/libbf_cuda.so > gpurun_out/ab_delta_inkernel.txt 2>&1
cat gpurun_out/ab_delta_inkernel.txt
timeout 400 python bench.py --cpu-sample 0 > gpurun_out/bench_line.json 2> gpurun_out/bench.err
cat gpurun_out/bench_line.json | cut -c1-1500; tail -3 gpurun_out/bench.err
