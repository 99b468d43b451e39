#!/bin/bash
# Multi-GPU development call: bench.py on N ranks (NGPU), optionally both DP modes.  gpurun --gpus N -- 'NGPU=N bash tools/gpu_call_multi.sh'
mkdir -p gpurun_out
N=${NGPU:-2}
for mode in ${MODES:-sharded sharded_noverlap allreduce}; do
  dp=$mode; ov=1; if [ "$mode" = "sharded_noverlap" ]; then dp=sharded; ov=0; fi
  PERF_B200_DP=$dp PERF_B200_AG_OVERLAP=$ov timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err; echo "bench n=$N $mode exit=$?"
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_n${N}_$mode.json'))
    print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1))
    print(json.dumps({k: v for k, v in (d.get('train') or {}).items() if k != 'note'}, indent=0))
    for k in ('render_c4', 'render_c5'):
        print(k, d.get(k))
except Exception as e:
    print('bench unreadable', e); print(open('gpurun_out/bench_n${N}_$mode.err').read()[-3000:])
PY
done
