#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider --durations=6 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" | tee gpurun_out/summary.txt
tail -14 gpurun_out/pytest_gpu.log
timeout 300 python tools/quick_bench.py > gpurun_out/quick_bench.log 2>&1; echo "quick_bench exit=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/quick_bench.log
