#!/bin/bash
# usage: gpu_r2_multi.sh N   (on an N-GPU lease): weak-scaling bench line (cfg2) and strong-scaling lines (cfg4, cfg5)
N=$1
mkdir -p gpurun_out
run() { # name, extra args
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/bench_${name}_${N}gpu.json 2> gpurun_out/bench_${name}_${N}gpu.err
  echo "== $name N=$N"; cat gpurun_out/bench_${name}_${N}gpu.json; tail -2 gpurun_out/bench_${name}_${N}gpu.err
}
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks_${N}gpu.csv &
SMI=$!
run weak_cfg2
run strong_cfg4 --config cfg4 --scaling strong
run strong_cfg5 --config cfg5 --scaling strong
kill $SMI
if [ "$N" = "2" ]; then
  ( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/pytest_multi_2gpu.log 2>&1; tail -3 gpurun_out/pytest_multi_2gpu.log
fi
