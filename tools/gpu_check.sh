#!/bin/bash
# Everything the round-end driver does, in one gpurun call:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_check.sh'            (1 GPU)
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'NGPU=8 bash tools/gpu_check.sh'
# Outputs land in gpurun_out/ (copy what should be judged into profiles/).
mkdir -p gpurun_out
N=${NGPU:-1}
if [ "$N" = "1" ]; then
  timeout 600 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider --timeout=200 --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" | tee gpurun_out/summary.txt
  tail -4 gpurun_out/pytest_gpu.log
  timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
  timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?" | tee -a gpurun_out/summary.txt
  timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref exit=$?" | tee -a gpurun_out/summary.txt
  if [ -n "$NCU" ]; then
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-train > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu-launches exit=$?" | tee -a gpurun_out/summary.txt
    ROWS=1024 timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_march -s 1 -c 1 -f -o gpurun_out/prof_bench_kernel python tools/prof_render.py > gpurun_out/prof_bench_kernel.log 2>&1; echo "ncu-full exit=$?" | tee -a gpurun_out/summary.txt
  fi
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n=$N exit=$?" | tee gpurun_out/summary.txt
fi
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/bench_n*.json')):
    try:
        d = json.load(open(f)); print(f, 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'frac', round(d['roofline']['frac'], 3), 'train', {k: round(v, 2) for k, v in (d.get('train') or {}).items() if k.endswith('ms_per_step')})
    except Exception as e:
        print(f, 'unreadable', e)
PY
