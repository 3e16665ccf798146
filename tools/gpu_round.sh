#!/bin/bash
# One GPU-box visit: tests, bench (both arms), launch list, full-set capture, phase counters.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
timeout 300 python bench.py > gpurun_out/bench_line.json 2> gpurun_out/bench.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_line_reference_arm.json 2> gpurun_out/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launch_list.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bf_minimize -s 1 -c 1 -f -o gpurun_out/minimize_full python tools/prof_batch.py 0.03 -1 592 0 1 > gpurun_out/ncu_full.log 2>&1
BF_PROFILE=1 timeout 200 python tools/prof_batch.py 0.03 -1 592 0 5 > gpurun_out/phase_cycles.txt 2>&1
timeout 300 python tools/bench_configs.py > gpurun_out/baseline_configs.txt 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_line.json; cat gpurun_out/phase_cycles.txt | head -3
