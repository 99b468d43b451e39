#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_train.py tests/test_gpu_basic.py -q -m gpu --no-header -p no:cacheprovider --durations=5 --timeout=150 -x > gpurun_out/pytest_train.log 2>&1; echo "pytest exit=$?" | tee gpurun_out/summary.txt
tail -12 gpurun_out/pytest_train.log
FUSED=1 timeout 200 python tools/train_bench.py > gpurun_out/train_bench.log 2>&1; echo "train_bench exit=$?" | tee -a gpurun_out/summary.txt
tail -4 gpurun_out/train_bench.log
FUSED=1 PHASES=geo NSTEPS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/train_launches.csv python tools/train_bench.py > gpurun_out/train_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/train_launches.csv')) if len(r) > 14 and r[0].isdigit()]
n = len(rows); tail = rows[int(n*0.72):]
agg = collections.OrderedDict()
for r in tail:
    k = r[4][:90]; agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += float(r[14])/1e3
print("launches in window", len(tail), "total us", sum(v[1] for v in agg.values()))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{v[1]:10.1f} us  x{v[0]:3d}  {k}")
PY
