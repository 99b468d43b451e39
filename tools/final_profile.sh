#!/bin/bash
# Round-end evidence in one call (1 GPU): launch lists + ncu --set full digests of every kernel class, microbenchmarks with
# their counters, sanitizers on the small configuration.  Everything lands under gpurun_out/final/ as text (reports are
# summarised on the box; only the render kernel's report is kept).
mkdir -p gpurun_out/final
O=gpurun_out/final
digest() { python tools/ncu_summary.py $O/$1.ncu-rep > $O/$1_summary.txt 2>&1; [ -z "$2" ] && rm -f $O/$1.ncu-rep; }
# 1. launch lists (cold-cache, serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-train > $O/bench_under_ncu.log 2>&1; echo "bench launches exit=$?"
for ph in geo app; do
  PHASES=$ph NSTEPS=2 GRAPH=0 FUSED=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/train_${ph}_launches.csv python tools/train_bench.py > $O/train_${ph}_ncu.log 2>&1; echo "train $ph launches exit=$?"
done
# 2. ncu --set full: render (bench launch), training kernels (both phases), packed path, microbenchmarks
ROWS=1024 timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_march -s 1 -c 1 -f -o $O/render_1024rows python tools/prof_render.py > /dev/null 2>&1; echo "ncu render exit=$?"; digest render_1024rows keep
[ $(stat -c %s $O/render_1024rows.ncu-rep 2>/dev/null || echo 0) -gt 45000000 ] && rm -f $O/render_1024rows.ncu-rep
for ph in geo app; do
  PHASES=$ph NSTEPS=1 GRAPH=0 timeout 600 ncu --set full --clock-control none -k regex:"render_march_kernel|composite_bwd|mlp_bwd_kernel|hashgrid_bwd|adam_kernel|train_loss|gather_rows" -s 40 -c 8 -f -o $O/train_$ph python tools/train_bench.py > /dev/null 2>&1; echo "ncu train $ph exit=$?"; digest train_$ph
done
OCC=1 PHASES=geo NSTEPS=1 GRAPH=0 timeout 600 ncu --set full --clock-control none -k regex:"packed_fields|composite_packed|hashgrid_bwd_kernel|occ_march" -s 30 -c 6 -f -o $O/train_occ_geo python tools/train_bench.py > $O/train_occ.log 2>&1; echo "ncu occ exit=$?"; digest train_occ_geo
timeout 600 ncu --set full --clock-control none -k regex:"hashgrid_fwd_kernel|network_fwd_kernel" -s 24 -c 8 -f -o $O/microbench python tools/encode_microbench.py > $O/microbench_under_ncu.log 2>&1; echo "ncu microbench exit=$?"; digest microbench
timeout 120 python tools/encode_microbench.py 2>&1 | tail -4 > $O/microbench.log
# 2b. end-to-end fit episodes with the reference's schedule (3000 + 1500 iterations), both samplers
timeout 300 python examples/fit_and_render.py 2>&1 | tail -3 > $O/fit_fixed.log; cat $O/fit_fixed.log
timeout 300 python examples/fit_and_render.py --sampler occ 2>&1 | tail -3 > $O/fit_occ.log; cat $O/fit_occ.log
# 3. sanitizers
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $O/compute_sanitizer_memcheck.log 2>&1; tail -2 $O/compute_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py > $O/compute_sanitizer_racecheck.log 2>&1; tail -2 $O/compute_sanitizer_racecheck.log
du -sh gpurun_out
