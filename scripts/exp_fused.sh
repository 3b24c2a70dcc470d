#!/bin/bash
# Round-2 experiment: shapes of the fused K1+K2 kernel (fast build under _variants/) by batch size.
#   gpurun --timeout 900 -- 'bash scripts/exp_fused.sh > gpurun_out/exp_fused.log 2>&1'
export QG_SOLVE_ONLY=1 ILQR_B200_LIB=$PWD/_variants/libfast.so
run() { echo "=== B=$1 $2"; env $2 QG_ITERS=${3:-10} python scripts/quick_gpu.py $1 500 rk4 | tail -2; }
for np in 2 3 4; do run 4096 "ILQR_FUSED_NP=$np ILQR_FUSED_MINB=1"; done
run 4096 "ILQR_FUSED=0"
for cfg in "ILQR_FUSED_NP=4" "ILQR_FUSED_NP=2 ILQR_FUSED_MINB=1" "ILQR_FUSED_NP=2 ILQR_FUSED_MINB=4"; do run 8192 "$cfg"; done
for cfg in "ILQR_FUSED_NP=2 ILQR_FUSED_MINB=1" "ILQR_FUSED_NP=2 ILQR_FUSED_MINB=4" "ILQR_FUSED_NP=2 ILQR_FUSED_MINB=5" "ILQR_FUSED=0"; do run 16384 "$cfg" 6; done
for cfg in "ILQR_FUSED_NP=2 ILQR_FUSED_MINB=4" "ILQR_FUSED_NP=2 ILQR_FUSED_MINB=5"; do run 32768 "$cfg" 6; done
