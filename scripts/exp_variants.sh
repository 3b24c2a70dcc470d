#!/bin/bash
# build macro variants of the library on the GPU box and time solves with each (exploration only)
#   scripts/exp_variants.sh "B N integ" "-DFLAG=1" "-DOTHER=2" ...
cd "$(dirname "$0")/.."
PK=iterative-linear-quadratic-regulator_b200
ARGS=$1; shift
for v in "$@"; do
  name=$(echo "$v" | tr -c 'A-Za-z0-9=\n' '_')
  out=/tmp/libilqr_$name.so
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Iinclude -I$PK/csrc $v -o $out $PK/csrc/ilqr_b200.cu || exit 1
  echo "== $v ($ARGS)"
  ILQR_B200_LIB=$out QG_SOLVE_ONLY=1 python scripts/quick_gpu.py $ARGS | tail -2
done
