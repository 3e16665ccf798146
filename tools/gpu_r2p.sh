#!/bin/bash
N=$1
mkdir -p gpurun_out
for rep in 1 2; do
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$rep bench.py --gpus $N --steps 5 --warmup 3 --config cfg5 --scaling strong > gpurun_out/bench_strong_cfg5_${N}gpu.json 2> gpurun_out/bench_strong_cfg5_${N}gpu.err
echo "rep $rep exit $?"
python -c "
import json
d=json.load(open('gpurun_out/bench_strong_cfg5_${N}gpu.json'))
print('cfg5 strong N=$N: value %.0f ms %.2f | e2e %.0f ms %.2f fmt %s | gather_ok %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['upload_format'], d.get('gather_ok')))
" || tail -3 gpurun_out/bench_strong_cfg5_${N}gpu.err | cut -c1-300
done
