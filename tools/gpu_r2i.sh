#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/ab_libs.py 592 2 better_flow_b200/libbf_cuda.so build/libbf_rcp512.so better_flow_b200/libbf_cuda.so:ctas_per_sm=1 better_flow_b200/libbf_cuda.so:ctas_per_sm=1,group_size=2 better_flow_b200/libbf_cuda.so:ctas_per_sm=1,group_size=1 > gpurun_out/ab_misc.txt 2>&1
cat gpurun_out/ab_misc.txt
( timeout 600 python -m pytest tests/test_gpu_cli.py -m gpu -q -k img ) 2>&1 | tail -2
