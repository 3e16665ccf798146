#!/bin/bash
# bench.py on every BASELINE configuration that fits one GPU (one JSON line each).
mkdir -p gpurun_out
for c in cfg2 cfg3 cfg4 cfg5; do
  timeout 400 python bench.py --config $c --steps 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  cut -c1-330 gpurun_out/bench_$c.json; tail -2 gpurun_out/bench_$c.err
done
timeout 300 python bench.py --config cfg4 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_cfg4_ref.json 2>&1; cut -c1-300 gpurun_out/bench_cfg4_ref.json
