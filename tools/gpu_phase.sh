#!/bin/bash
# Phase counters (in-kernel clock64) of the listed builds on the bench workload.
mkdir -p gpurun_out
for l in $AB_LIBS; do
  echo "== $l"; BF_LIB_PATH=$PWD/$l BF_PROFILE=1 timeout 200 python tools/prof_batch.py 0.03 -1 592 0 3 2>&1 | head -12
done > gpurun_out/phase_ab.txt 2>&1
cat gpurun_out/phase_ab.txt
