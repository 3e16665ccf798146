#!/bin/bash
N=$1
mkdir -p gpurun_out
for fmt in plain delta; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --upload-format $fmt > gpurun_out/bench_weak_cfg2_${N}gpu_${fmt}.json 2> gpurun_out/bench_weak_${N}gpu_${fmt}.err
  echo "== $fmt N=$N"; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_weak_cfg2_${N}gpu_${fmt}.json'))
print('value %.0f ms %.2f | e2e %.0f ms %.2f fmt %s h2d %d gather_ok %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['upload_format'], d['e2e']['h2d_bytes_per_step'], d.get('gather_ok')))
print([round(p['own_ms_per_step'],2) for p in d['per_rank']], [round(p['own_e2e_ms_per_step'],2) for p in d['per_rank']])
" || tail -5 gpurun_out/bench_weak_${N}gpu_${fmt}.err
done
