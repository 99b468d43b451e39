#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n=$N exit=$?" | tee gpurun_out/summary.txt
cat gpurun_out/bench_n$N.json | cut -c1-600; tail -3 gpurun_out/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/train_bench.py > gpurun_out/train_n$N.log 2>&1; echo "train n=$N exit=$?" | tee -a gpurun_out/summary.txt
grep "train step" gpurun_out/train_n$N.log; tail -3 gpurun_out/train_n$N.log
