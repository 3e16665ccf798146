#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ring.py -m gpu -x -q -k "cluster or ring" ) 2>&1 | tail -15
timeout 600 python tools/cli_ring.py 1.5 2>&1 | tee gpurun_out/cli_ring.txt
timeout 600 python tools/ab_libs.py 592 2 better_flow_b200/libbf_cuda.so better_flow_b200/libbf_cuda.so:cluster=2 better_flow_b200/libbf_cuda.so:cluster=4 > gpurun_out/ab_cluster.txt 2>&1
cat gpurun_out/ab_cluster.txt
