#!/bin/bash
# The three LTV Riccati kernels side by side on the GPU box: sixteen lanes per trajectory, four lanes x four columns,
# FP64 tensor cores (one warp per trajectory), at several batch sizes.   scripts/exp_ltv.sh [N]
cd "$(dirname "$0")/.."
N=${1:-1000}
for B in 256 1024 4096 32768; do
  for lanes in 16 4 32; do
    echo "== ILQR_LTV_LANES=$lanes  B=$B N=$N"
    ILQR_LTV_LANES=$lanes QG_ITERS=2 QG_ALPHAS=4 python scripts/quick_gpu.py $B $N ltv | tail -1
  done
done
