#!/bin/bash
mkdir -p gpurun_out
BF_PROFILE_PHASES=1 timeout 200 python tools/ring_latency.py 40 > gpurun_out/ring_latency_r2x.txt 2>&1
cat gpurun_out/ring_latency_r2x.txt
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file gpurun_out/ring_latency_ncu.csv python tools/ring_latency.py 25 > gpurun_out/ring_latency_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/ring_latency_ncu.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
d = collections.defaultdict(list)
for r in rows[1:]:
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e3 if u in ('ns', 'nsecond') else v
    d[r[ki][:60]].append(v)
for k, v in d.items():
    v = v[5:] if len(v) > 10 else v
    print("%-62s n %3d  mean %.1f us  min %.1f  max %.1f" % (k, len(v), sum(v) / len(v), min(v), max(v)))
PY
