#!/bin/bash
mkdir -p gpurun_out
OCC=1 PHASES=${PHASES:-geo} NSTEPS=2 GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/occ_launches.csv python tools/train_bench.py > gpurun_out/occ_ncu.log 2>&1
tail -3 gpurun_out/occ_ncu.log
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/occ_launches.csv')) if len(r) > 14 and r[0].isdigit()]
adam = [i for i, r in enumerate(rows) if 'adam_kernel' in r[4]]
lo, hi = adam[-2] + 1, adam[-1] + 1
agg = collections.OrderedDict()
for r in rows[lo:hi]:
    k = r[4][:100]; agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += float(r[14]) / 1e3
print("one occ step:", hi - lo, "launches,", round(sum(v[1] for v in agg.values()), 1), "us (cold-cache, serialised)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"{v[1]:9.1f} us  x{v[0]:2d}  {k}")
PY

