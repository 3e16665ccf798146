#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ab_libs.py 592 2 better_flow_b200/libbf_cuda.so better_flow_b200/libbf_cuda.so:smem_pad=8192 better_flow_b200/libbf_cuda.so:smem_pad=24576 better_flow_b200/libbf_cuda.so:smem_pad=49152 build/libbf_smallsmem.so > gpurun_out/ab_smem.txt 2>&1
cat gpurun_out/ab_smem.txt
timeout 600 python tools/cli_ring.py 1.5 ref > gpurun_out/cli_ring.txt 2>&1; cat gpurun_out/cli_ring.txt
