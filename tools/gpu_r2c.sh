#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ring.py tests/test_gpu_cli.py -m gpu -x -q -s ) > gpurun_out/pytest_ring.log 2>&1
tail -6 gpurun_out/pytest_ring.log
timeout 600 python tools/cli_ring.py 1.5 > gpurun_out/cli_ring.txt 2>&1; cat gpurun_out/cli_ring.txt
bash tools/gpu_strict.sh
timeout 600 python tools/ab_libs.py 592 2 better_flow_b200/libbf_cuda.so:group_size=2 better_flow_b200/libbf_cuda.so:group_size=3 better_flow_b200/libbf_cuda.so:group_size=4 better_flow_b200/libbf_cuda.so:group_size=5 better_flow_b200/libbf_cuda.so:group_size=6 > gpurun_out/ab_groups.txt 2>&1
cat gpurun_out/ab_groups.txt
for G in 2 4 16; do timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:bf_minimize -s 1 -c 1 python tools/prof_batch.py 0.03 -1 592 $G 1 2>&1 | grep -E "dram__|lts__|gpu__time|slice " ; done > gpurun_out/ncu_groups_dram.txt 2>&1
cat gpurun_out/ncu_groups_dram.txt
