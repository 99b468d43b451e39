#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_train.py -q -m gpu --no-header -p no:cacheprovider --durations=6 --timeout=150 > gpurun_out/pytest_train.log 2>&1; echo "pytest-train exit=$?" | tee gpurun_out/summary.txt
tail -12 gpurun_out/pytest_train.log
timeout 200 python tools/train_bench.py > gpurun_out/train_bench.log 2>&1; echo "train_bench exit=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/train_bench.log
