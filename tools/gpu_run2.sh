#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" | tee gpurun_out/summary.txt
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref exit=$?" | tee -a gpurun_out/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu-launches exit=$?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/prof_render python tools/prof_render.py > gpurun_out/prof_render.log 2>&1; echo "ncu-full exit=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/prof_render.log
python __graft_entry__.py smoke 2>&1 | tail -2
