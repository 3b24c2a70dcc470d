set -x
python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1f_ref.json 2> gpurun_out/bench_r1f_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --large-batch 0 > gpurun_out/ncu1f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rollout|linearize|backward" -s 12 -c 6 -o gpurun_out/prof_r1f -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --large-batch 0 > gpurun_out/ncu2f.log 2>&1
QG_SOLVE_ONLY=1 QG_NO_REPS=1 QG_ITERS=2 ncu --set full --clock-control none -k regex:"rollout|linearize|backward" -c 14 -o gpurun_out/prof_r1f_large -f python scripts/quick_gpu.py 131072 500 rk4 > gpurun_out/ncu3f.log 2>&1
python scripts/bench_configs.py > gpurun_out/configs_r1f.json 2> gpurun_out/configs_r1f.err
ls -la gpurun_out/*r1f*
