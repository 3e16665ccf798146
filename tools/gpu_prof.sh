#!/bin/bash
# Profile visit: bench line, launch list, full-set capture and phase counters of the current build.
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/bench_line.json 2> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launch_list.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bf_minimize -s 1 -c 1 -f -o gpurun_out/minimize_full python tools/prof_batch.py 0.03 -1 592 0 1 > gpurun_out/ncu_full.log 2>&1
BF_PROFILE=1 timeout 200 python tools/prof_batch.py 0.03 -1 592 0 5 > gpurun_out/phase_cycles.txt 2>&1
cat gpurun_out/bench_line.json; head -4 gpurun_out/phase_cycles.txt
