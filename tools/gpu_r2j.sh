#!/bin/bash
# Round 2, full visit: every GPU test, the bench lines (both arms), launch list, full-set capture, phase counters.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_line.json 2> gpurun_out/bench.err
cat gpurun_out/bench_line.json; tail -2 gpurun_out/bench.err
timeout 400 python bench.py --upload-format plain --cpu-sample 0 > gpurun_out/bench_line_plain_upload.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_line_reference_arm.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_line_reference_arm.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launch_list.csv python bench.py --steps 2 --warmup 3 --cpu-sample 0 > gpurun_out/ncu_bench.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:bf_minimize -s 1 -c 1 -f -o gpurun_out/minimize_full python tools/prof_batch.py 0.03 -1 592 0 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
BF_PROFILE=1 timeout 200 python tools/prof_batch.py 0.03 -1 592 0 5 > gpurun_out/phase_cycles.txt 2>&1
