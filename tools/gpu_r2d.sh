#!/bin/bash
mkdir -p gpurun_out
( BF_LIB_PATH=$PWD/build/libbf_tma.so timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_tma.log 2>&1
tail -4 gpurun_out/pytest_tma.log
timeout 600 python tools/ab_libs.py 592 3 better_flow_b200/libbf_cuda.so build/libbf_tma.so build/libbf_tma.so:group_size=2 > gpurun_out/ab_tma.txt 2>&1
cat gpurun_out/ab_tma.txt
bash tools/gpu_strict.sh
for l in better_flow_b200/libbf_cuda.so build/libbf_tma.so; do echo "== $l"; BF_LIB_PATH=$PWD/$l BF_PROFILE=1 timeout 200 python tools/prof_batch.py 0.03 -1 592 0 3 2>&1 | head -20; done > gpurun_out/phase_tma.txt 2>&1
cat gpurun_out/phase_tma.txt
