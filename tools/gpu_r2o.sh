#!/bin/bash
N=$1
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/bench_${name}_${N}gpu.json 2> gpurun_out/bench_${name}_${N}gpu.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_${name}_${N}gpu.json'))
print('$name N=$N: value %.0f ms %.2f | e2e %.0f ms %.2f fmt %s | gather_ok %s parity %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['upload_format'], d.get('gather_ok'), d.get('parity',{}).get('max_rel_dxdy')))
print('   own ms', [round(p['own_ms_per_step'],2) for p in d['per_rank']], 'own e2e ms', [round(p['own_e2e_ms_per_step'],2) for p in d['per_rank']], 'iters', [p['sum_iters'] for p in d['per_rank']])
" || tail -5 gpurun_out/bench_${name}_${N}gpu.err
}
run weak_cfg2
run strong_cfg4 --config cfg4 --scaling strong
run strong_cfg5 --config cfg5 --scaling strong
