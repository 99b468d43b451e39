#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_train.py -q -m gpu --no-header -p no:cacheprovider --timeout=150 -x -s -k "fit or fused_train" > gpurun_out/pytest_train.log 2>&1; echo "pytest exit=$?" | tee gpurun_out/summary.txt
grep -E "PSNR|passed|failed" gpurun_out/pytest_train.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" | tee -a gpurun_out/summary.txt
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac']); print(d.get('train')); print(d.get('cpu_baseline'))"
tail -2 gpurun_out/bench.err
