#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider --timeout=150 --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" | tee gpurun_out/summary.txt
tail -14 gpurun_out/pytest_gpu.log
FUSED=1 timeout 200 python tools/train_bench.py > gpurun_out/train_bench.log 2>&1; echo "train_bench exit=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/train_bench.log
