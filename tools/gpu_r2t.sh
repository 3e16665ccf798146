#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
for cfg in cfg3 cfg4 cfg5; do
  timeout 300 python bench.py --config $cfg --cpu-sample 4 > gpurun_out/bench_line_$cfg.json 2> gpurun_out/bench_$cfg.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_line_$cfg.json'))
print('$cfg: value %.0f ms %.3f | e2e %.0f ms %.3f | frac %.3f | parity %s | cpu %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d.get('parity',{}).get('max_rel_dxdy'), d.get('cpu_baseline',{}).get('value')))
" || tail -3 gpurun_out/bench_$cfg.err
done
timeout 300 python tools/bench_configs.py > gpurun_out/baseline_configs.txt 2>&1; cat gpurun_out/baseline_configs.txt
