#!/bin/bash
# Round-end evidence after the instruction-count / cell-major / scattered-tile work on the field kernels (one GPU):
# full GPU suite, both bench arms, ncu --set full digest of the bench's render launch (+ the counters bench.py quotes),
# launch lists of the bench and of one training step.
O=gpurun_out/final3; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider --timeout=300 > $O/pytest_gpu.log 2>&1; echo "pytest exit=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $O/pytest_gpu.log | tail -10
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench exit=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref exit=$?"
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/final3/bench_n1.json'))
    print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'parity', d.get('parity', {}).get('max_abs_rgb'))
    t = d.get('train') or {}
    print('train', {k: round(v, 3) for k, v in t.items() if isinstance(v, float)}, 'occ', {k: (round(v, 3) if isinstance(v, float) else v) for k, v in (t.get('occ') or {}).items() if 'ms' in k})
    for k in ('render_c4', 'render_c5'):
        print(k, (d.get(k) or {}).get('msamples_per_s'))
    print('cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('cores'))
    r = json.load(open('gpurun_out/final3/bench_ref.json')); print('ref', r.get('value'), r.get('unit'))
except Exception as e:
    print('bench unreadable', e); print(open('gpurun_out/final3/bench_n1.err').read()[-3000:])
PY
ROWS=1024 timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_march -s 1 -c 1 -f -o $O/render_1024rows python tools/prof_render.py > /dev/null 2>&1; echo "ncu render exit=$?"
python tools/ncu_summary.py $O/render_1024rows.ncu-rep > $O/render_1024rows_summary.txt 2>&1
ncu -i $O/render_1024rows.ncu-rep --page source --csv > $O/render_1024rows_source.csv 2>/dev/null
rm -f $O/render_1024rows.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-train > $O/bench_under_ncu.log 2>&1; echo "bench launches exit=$?"
PHASES=geo NSTEPS=2 GRAPH=0 FUSED=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/train_geo_launches.csv python tools/train_bench.py > $O/train_geo_ncu.log 2>&1; echo "train launches exit=$?"
python - <<'PY' | tee gpurun_out/final3/launch_summary.txt
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/final3/train_geo_launches.csv')) if len(r) > 14 and r[0].isdigit()]
adam = [i for i, r in enumerate(rows) if 'adam_kernel' in r[4]]
lo, hi = adam[-2] + 1, adam[-1] + 1
agg = collections.OrderedDict()
for r in rows[lo:hi]:
    k = r[4][:90]; agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += float(r[14]) / 1e3
print("one density step:", hi - lo, "launches,", round(sum(v[1] for v in agg.values()), 1), "us (cold-cache, serialised)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{v[1]:9.1f} us  x{v[0]:2d}  {k}")
PY
du -sh $O
