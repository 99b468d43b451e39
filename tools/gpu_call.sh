#!/bin/bash
# Generic development call: GPU tests (all, no -x), bench line, training launch lists, one ncu --set full of the render kernel.
# Outputs under gpurun_out/.  Env: PYTEST_K (subset), SKIP_NCU=1, SKIP_TESTS=1.
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  timeout 1200 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider --timeout=300 ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"
  grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -25
fi
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_n1.json'))
    print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'frac', round(d['roofline']['frac'], 3), 'parity', d.get('parity', {}).get('max_abs_rgb'))
    print('train', {k: (round(v, 3) if isinstance(v, float) else v) for k, v in (d.get('train') or {}).items() if k != 'note'})
    for k in ('render_c4', 'render_c5'):
        print(k, d.get(k))
    print('cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('cores'))
except Exception as e:
    print('bench unreadable', e); print(open('gpurun_out/bench_n1.err').read()[-3000:])
PY
timeout 300 python tools/render_variants.py 5 2>&1 | tail -5 | tee gpurun_out/render_variants.log
timeout 120 python -c "
import sys; sys.path.insert(0, '.')
from perf_b200 import ops
for vec in (4, 2, 1):
    print('L2 atomic rate, random %d-float reductions into the 16.8 MB fine-level gradient: %.1f G atomics/s' % (vec, ops.atomic_rate(vec=vec) / 1e9))
" 2>&1 | tail -3 | tee gpurun_out/atomic_rate.log
timeout 120 python tools/ab_mlp_bwd.py 2>&1 | tail -2 | tee gpurun_out/ab_mlp_bwd.log
timeout 120 python tools/ab_scatter_v4.py 2>&1 | tail -3 | tee gpurun_out/ab_scatter.log
if [ -z "$SKIP_NCU" ]; then
  PHASES=geo bash tools/profile_train.sh 2>&1 | tail -18; cp gpurun_out/train_launches.csv gpurun_out/train_launches_geo.csv
  PHASES=app bash tools/profile_train.sh 2>&1 | tail -18; cp gpurun_out/train_launches.csv gpurun_out/train_launches_app.csv
  digest() {   # summarise on the box, keep the report only if it is small (gpurun brings back <= 64 MiB)
    python tools/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/$1_summary.txt 2>&1
    if [ $(stat -c %s gpurun_out/$1.ncu-rep 2>/dev/null || echo 0) -gt ${KEEP_REP_BYTES:-12000000} ]; then rm -f gpurun_out/$1.ncu-rep; fi
  }
  for ph in geo app; do
    PHASES=$ph NSTEPS=1 GRAPH=0 timeout 600 ncu --set full --clock-control none -k regex:"render_march_kernel|composite_bwd|mlp_bwd_kernel|hashgrid_bwd|adam_kernel" -s 30 -c 6 -f -o gpurun_out/prof_train_${ph}_r02 python tools/train_bench.py > gpurun_out/prof_train_${ph}_r02.log 2>&1; echo "ncu-train-$ph exit=$?"
    digest prof_train_${ph}_r02
  done
  ROWS=256 timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_march -s 1 -c 1 -f -o gpurun_out/prof_render_r02 python tools/prof_render.py > gpurun_out/prof_render_r02.log 2>&1; echo "ncu-full exit=$?"
  KEEP_REP_BYTES=40000000 digest prof_render_r02
  KERNEL=march_l0smem ROWS=256 timeout 600 ncu --set full --clock-control none -k regex:render_march -s 1 -c 1 -f -o gpurun_out/prof_render_l0smem_r02 python tools/prof_render.py > gpurun_out/prof_render_l0smem_r02.log 2>&1; echo "ncu-full-l0 exit=$?"
  digest prof_render_l0smem_r02
  ROWS=1024 timeout 600 ncu --set full --clock-control none -k regex:render_march -s 1 -c 1 -f -o gpurun_out/prof_render_1024rows_r02 python tools/prof_render.py > /dev/null 2>&1; echo "ncu-full-1024 exit=$?"
  digest prof_render_1024rows_r02
  du -sh gpurun_out
fi
