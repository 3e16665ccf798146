#!/bin/bash
# One full-set capture of the minimise kernel of the current build on the bench workload.
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bf_minimize -s 1 -c 1 -f -o gpurun_out/minimize_full python tools/prof_batch.py 0.03 -1 592 0 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
