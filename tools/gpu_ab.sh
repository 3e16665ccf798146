#!/bin/bash
# A/B visit: parity tests with the current build, then interleaved timing of option variants, then a full-set capture.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python tools/ab_libs.py 592 3 $AB_LIBS > gpurun_out/ab.txt 2>&1
cat gpurun_out/ab.txt
if [ -n "$AB_NCU" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bf_minimize -s 1 -c 1 -f -o gpurun_out/minimize_full python tools/prof_batch.py 0.03 -1 592 0 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
fi
