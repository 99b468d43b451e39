#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" | tee gpurun_out/summary.txt
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu-launches exit=$?" | tee -a gpurun_out/summary.txt
ROWS=1024 timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_march -s 1 -c 1 -f -o gpurun_out/prof_bench_kernel python tools/prof_render.py > gpurun_out/prof_bench_kernel.log 2>&1; echo "ncu-full exit=$?" | tee -a gpurun_out/summary.txt
tail -2 gpurun_out/prof_bench_kernel.log
