#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_network.py -q -m gpu --no-header -p no:cacheprovider --durations=8  > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" | tee gpurun_out/summary.txt
tail -16 gpurun_out/pytest_gpu.log
timeout 300 python tools/quick_bench.py > gpurun_out/quick_bench.log 2>&1; echo "quick_bench exit=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/quick_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_march -s 1 -c 1 -f -o gpurun_out/prof_march python tools/prof_render.py > gpurun_out/prof_march.log 2>&1; echo "ncu-full exit=$?" | tee -a gpurun_out/summary.txt
tail -2 gpurun_out/prof_march.log
