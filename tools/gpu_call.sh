#!/bin/bash
# Generic development call: GPU tests, bench line, graphed training-step timing.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider --timeout=300 -x ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_n1.json'))
    print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'frac', round(d['roofline']['frac'], 3), 'parity', d.get('parity', {}).get('max_abs_rgb'))
    print('train', {k: round(v, 3) for k, v in (d.get('train') or {}).items() if isinstance(v, float)})
except Exception as e:
    print('bench unreadable', e); print(open('gpurun_out/bench_n1.err').read()[-3000:])
PY
