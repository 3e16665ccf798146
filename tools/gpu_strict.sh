#!/bin/bash
# Validation build of the group barrier (ld.acquire + fence.acq_rel instead of the relaxed poll): the GPU parity
# tests must pass unchanged with it.  usage (on the GPU box): bash tools/gpu_strict.sh
mkdir -p gpurun_out
( BF_LIB_PATH=$PWD/build/libbf_strict.so timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_baseline_sizes.py -m gpu -x -q ) > gpurun_out/pytest_strict_sync.log 2>&1
tail -3 gpurun_out/pytest_strict_sync.log
