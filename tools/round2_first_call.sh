#!/bin/bash
# First GPU call of the next round: everything written at the end of round 1 without GPU time.
#   /usr/local/graft/bin/gpurun --timeout 400 -- 'bash tools/round2_first_call.sh'
# Each step has its own timeout: the experimental kernel traps (bounded mbarrier wait) rather than hangs, but
# a trap poisons the CUDA context of that process only.
mkdir -p gpurun_out
export PERF_B200_EXPERIMENTAL=1
for k in "simt and density" "simt and colour" "tcgen05 and density" "tcgen05 and colour"; do
  timeout 90 python -m pytest tests/test_gpu_train.py -q -x --timeout 60 -k "single_kernel_mlp_backward and $k" 2>&1 | tail -4 | tee -a gpurun_out/round2_first.log
done
timeout 120 python tools/ab_mlp_bwd.py 2>&1 | tail -6 | tee -a gpurun_out/round2_first.log
timeout 60 python tools/ab_scatter_v4.py 2>&1 | tail -2 | tee -a gpurun_out/round2_first.log
timeout 120 python tools/encode_microbench.py 2>&1 | tail -5 | tee -a gpurun_out/round2_first.log
