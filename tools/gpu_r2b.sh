#!/bin/bash
mkdir -p gpurun_out
./build/cluster_query > gpurun_out/cluster_query.txt 2>&1; cat gpurun_out/cluster_query.txt
timeout 900 python tools/ab_libs.py 592 2 better_flow_b200/libbf_cuda.so build/libbf_stampb.so better_flow_b200/libbf_cuda.so:group_size=4 better_flow_b200/libbf_cuda.so:group_size=8 better_flow_b200/libbf_cuda.so:group_size=16 better_flow_b200/libbf_cuda.so:group_size=16,tail_help=0 > gpurun_out/ab.txt 2>&1
cat gpurun_out/ab.txt
for G in 2 8 16; do echo "== G $G"; BF_PROFILE=1 timeout 200 python tools/prof_batch.py 0.03 -1 592 $G 3 2>&1 | head -20; done > gpurun_out/phase_groups.txt 2>&1
cat gpurun_out/phase_groups.txt
