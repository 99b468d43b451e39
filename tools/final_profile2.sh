#!/bin/bash
# Last evidence call of the round: launch lists + ncu digests of the training kernels as they are at round end (MLP backward
# with the fine scatter in its epilogue, coarse march), then the bench line.
mkdir -p gpurun_out/final2
O=gpurun_out/final2
digest() { python tools/ncu_summary.py $O/$1.ncu-rep > $O/$1_summary.txt 2>&1; rm -f $O/$1.ncu-rep; }
for ph in geo app; do
  PHASES=$ph NSTEPS=2 GRAPH=0 FUSED=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/train_${ph}_launches.csv python tools/train_bench.py > $O/train_${ph}_ncu.log 2>&1; echo "train $ph launches exit=$?"
  python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('$O/train_${ph}_launches.csv')) if len(r) > 14 and r[0].isdigit()]
adam = [i for i, r in enumerate(rows) if 'adam_kernel' in r[4]]
lo, hi = adam[-2] + 1, adam[-1] + 1
agg = collections.OrderedDict()
for r in rows[lo:hi]:
    k = r[4][:90]; agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += float(r[14]) / 1e3
print("one $ph step:", hi - lo, "launches,", round(sum(v[1] for v in agg.values()), 1), "us (cold-cache, serialised)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{v[1]:9.1f} us  x{v[0]:2d}  {k}")
PY
  PHASES=$ph NSTEPS=1 GRAPH=0 timeout 600 ncu --set full --clock-control none -k regex:"render_march_kernel|composite_bwd|mlp_bwd_kernel|hashgrid_bwd" -s 20 -c 4 -f -o $O/train_$ph python tools/train_bench.py > /dev/null 2>&1; echo "ncu train $ph exit=$?"; digest train_$ph
done 2>&1 | tee $O/launch_summary.txt
python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench exit=$?"
timeout 200 python examples/fit_and_render.py 2>&1 | tail -3 | tee $O/fit_fixed.log
du -sh gpurun_out
