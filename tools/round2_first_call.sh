#!/bin/bash
# First GPU call of round 2: validate csrc/mlp_bwd.cu (written blind in round 1), then the A/B and microbenchmarks.
mkdir -p gpurun_out
L=gpurun_out/round2_first.log; : > $L
for net in density colour; do for mode in simt tc; do
  timeout 60 python tools/diag_mlp_bwd.py $net $mode 0 2>&1 | tail -5 | tee -a $L
done; done
for net in density colour; do timeout 60 python tools/diag_mlp_bwd.py $net tc 1 2>&1 | tail -5 | tee -a $L; done
timeout 120 python tools/ab_mlp_bwd.py 2>&1 | tail -6 | tee -a $L
timeout 60 python tools/ab_scatter_v4.py 2>&1 | tail -2 | tee -a $L
timeout 120 python tools/encode_microbench.py 2>&1 | tail -5 | tee -a $L
GRAPH=1 timeout 120 python tools/train_bench.py 2>&1 | tail -3 | tee -a $L
PHASES=geo bash tools/profile_train.sh 2>&1 | tail -20 | tee -a $L
cp gpurun_out/train_launches.csv gpurun_out/train_launches_geo.csv
PHASES=app bash tools/profile_train.sh 2>&1 | tail -20 | tee -a $L
