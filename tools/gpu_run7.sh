#!/bin/bash
mkdir -p gpurun_out
PHASES=geo NSTEPS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/train_launches.csv python tools/train_bench.py > gpurun_out/train_ncu.log 2>&1; echo "ncu exit=$?" | tee gpurun_out/summary.txt
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/train_launches.csv')) if len(r) > 14 and r[0].isdigit()]
# last 2 steps: take the tail window after warmup: group by kernel name over the final 40% of launches
n = len(rows); tail = rows[int(n*0.72):]
agg = collections.OrderedDict()
for r in tail:
    k = r[4][:90]; agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += float(r[14])/1e3
tot = sum(v[1] for v in agg.values())
print("launches in window", len(tail), "total us", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{v[1]:10.1f} us  x{v[0]:3d}  {k}")
PY
