#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the 2-device tests, then bench.py under torchrun on N GPUs.
N=${1:-2}; shift
mkdir -p gpurun_out
if [ "$N" = "2" ]; then ( timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/pytest_multi.log 2>&1; tail -2 gpurun_out/pytest_multi.log; fi
for c in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config $c --steps 5 --warmup 3 > gpurun_out/bench_${c}_${N}gpu.json 2> gpurun_out/bench_${c}_${N}gpu.err
  cut -c1-260 gpurun_out/bench_${c}_${N}gpu.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_${c}_${N}gpu.json; tail -1 gpurun_out/bench_${c}_${N}gpu.err | cut -c1-200
done
