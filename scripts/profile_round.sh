#!/bin/bash
# One GPU call that refreshes the evidence under profiles/ for a round tag (default r02):
#   gpurun --timeout 1500 -- 'bash scripts/profile_round.sh r02'
# bench line, reference arm, ncu launch list, ncu --set full captures at B=4096 (kept: source view) and
# B=131072 (summarised on the box, the 60 MB report is not brought back), other configurations.
# PROFILE_SHORT=1 skips the reference arm and the other configurations.
tag=${1:-r02}
out=gpurun_out
set -x
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err
[ -n "$PROFILE_SHORT" ] || python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_${tag}_ref.json 2> $out/bench_${tag}_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --large-batch 0 > $out/ncu1_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rollout|linearize|backward" -s 12 -c 6 -o $out/prof_$tag -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --large-batch 0 > $out/ncu2_$tag.log 2>&1
python scripts/ncu_summary.py $out/prof_$tag.ncu-rep > $out/${tag}_ncu_full_B4096.txt
python scripts/ncu_summary.py $out/prof_$tag.ncu-rep --traffic B4096 $out/traffic_$tag.json
QG_SOLVE_ONLY=1 QG_NO_REPS=1 QG_ITERS=2 ncu --set full --clock-control none -k regex:"rollout|linearize|backward" -c 14 -o /tmp/prof_${tag}_large -f python scripts/quick_gpu.py 131072 500 rk4 > $out/ncu3_$tag.log 2>&1
python scripts/ncu_summary.py /tmp/prof_${tag}_large.ncu-rep > $out/${tag}_ncu_full_B131072.txt
python scripts/ncu_summary.py /tmp/prof_${tag}_large.ncu-rep --traffic B131072 $out/traffic_$tag.json
# the two-kernel path's Riccati scan at B=131072 (bulk-copy ring; the kernel of north_star's HBM target)
QG_SOLVE_ONLY=1 QG_NO_REPS=1 QG_ITERS=2 ILQR_FUSED=0 ncu --set full --clock-control none -k regex:"backward_kernel" -c 2 -o /tmp/prof_${tag}_k2 -f python scripts/quick_gpu.py 131072 500 rk4 > $out/ncu5_$tag.log 2>&1
python scripts/ncu_summary.py /tmp/prof_${tag}_k2.ncu-rep > $out/${tag}_ncu_full_K2_B131072.txt
# config 4 (LTV, n=12, m=4): the sixteen-lane Riccati kernel and the n=12 rollout, incl. the shared-memory counters
QG_SOLVE_ONLY=1 QG_NO_REPS=1 QG_ITERS=1 QG_ALPHAS=4 ncu --set full --clock-control none --import-source on -k regex:"backward_ltv|rollout" -c 4 -o /tmp/prof_${tag}_ltv -f python scripts/quick_gpu.py 32768 1000 ltv > $out/ncu4_$tag.log 2>&1
python scripts/ncu_summary.py /tmp/prof_${tag}_ltv.ncu-rep > $out/${tag}_ncu_full_ltv_B32768.txt
[ -n "$PROFILE_SHORT" ] || python scripts/bench_configs.py > $out/configs_$tag.json 2> $out/configs_$tag.err
# only the B=4096 report (source view of the hot kernels) travels back: gpurun_out/ is limited to 64 MiB
ls -la $out/*$tag*; du -sh $out
