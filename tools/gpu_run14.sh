#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider --timeout=200 --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" | tee gpurun_out/summary.txt
tail -14 gpurun_out/pytest_gpu.log
