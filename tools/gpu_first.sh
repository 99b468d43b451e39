#!/bin/bash
# first bring-up on the GPU box: every stage in its own process, under a timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for t in test_gpu_basic test_gpu_network test_gpu_render; do
  timeout 600 python -m pytest tests/$t.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit=$?" | tee -a gpurun_out/summary.txt
  tail -5 gpurun_out/$t.log
done
timeout 300 python tools/quick_bench.py > gpurun_out/quick_bench.log 2>&1; echo "quick_bench exit=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/quick_bench.log
