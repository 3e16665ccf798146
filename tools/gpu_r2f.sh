#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ab_libs.py 592 2 better_flow_b200/libbf_cuda.so build/libbf_small.so build/libbf_tmafi.so build/libbf_tmasmall.so > gpurun_out/ab_tma2.txt 2>&1
cat gpurun_out/ab_tma2.txt
for l in better_flow_b200/libbf_cuda.so build/libbf_small.so build/libbf_tmasmall.so; do echo "== $l"; BF_LIB_PATH=$PWD/$l timeout 400 python tools/bench_configs.py 2>&1; done > gpurun_out/configs_ab.txt 2>&1
cat gpurun_out/configs_ab.txt
for l in build/libbf_tmasmall.so; do echo "== $l"; BF_LIB_PATH=$PWD/$l BF_PROFILE=1 timeout 200 python tools/prof_batch.py 0.03 -1 592 0 3 2>&1 | head -20; done > gpurun_out/phase_tma2.txt 2>&1
cat gpurun_out/phase_tma2.txt
