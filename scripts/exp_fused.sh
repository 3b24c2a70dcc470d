#!/bin/bash
# Round-2 experiment: fused K1+K2 kernel variants (fast builds under _variants/) against the two-kernel path.
#   gpurun --timeout 900 -- 'bash scripts/exp_fused.sh > gpurun_out/exp_fused.log 2>&1'
export QG_SOLVE_ONLY=1 QG_ITERS=${QG_ITERS:-4}
for B in 131072 4096; do
  echo "=== B=$B  two-kernel path (ILQR_FUSED=0)"
  ILQR_B200_LIB=$PWD/_variants/libfast_mb1.so ILQR_FUSED=0 python scripts/quick_gpu.py $B 500 rk4 | tail -3
  for v in mb1 mb4 mb5; do
    echo "=== B=$B  fused NP=2 $v"
    ILQR_B200_LIB=$PWD/_variants/libfast_$v.so ILQR_FUSED=1 python scripts/quick_gpu.py $B 500 rk4 | tail -3
  done
  echo "=== B=$B  fused NP=3 mb4"
  ILQR_B200_LIB=$PWD/_variants/libfast_mb4.so ILQR_FUSED=1 ILQR_FUSED_NP=3 python scripts/quick_gpu.py $B 500 rk4 | tail -3
done
