#!/bin/bash
# Build macro variants of the library HERE (nvcc cross-compiles) and time solves with each on the GPU box
# (exploration only).  Step 1, locally:   scripts/exp_variants.sh build "-DFLAG=1" "-DOTHER=2" ...
#                      Step 2, on the box: scripts/exp_variants.sh run "B N integ" ["B N integ" ...]
cd "$(dirname "$0")/.."
PK=iterative-linear-quadratic-regulator_b200
mode=$1; shift
if [ "$mode" = build ]; then
  rm -rf _variants; mkdir -p _variants
  for v in "$@"; do
    name=$(echo "$v" | tr -c 'A-Za-z0-9=\n' '_')
    ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Iinclude -I$PK/csrc $v \
        -o _variants/libilqr_$name.so $PK/csrc/ilqr_b200.cu 2>/dev/null || echo "FAILED $v" ) &
  done
  wait; ls _variants
else
  for a in "$@"; do
    for so in _variants/*.so; do
      echo "== $so ($a)"
      ILQR_B200_LIB=$PWD/$so QG_SOLVE_ONLY=1 python scripts/quick_gpu.py $a | tail -3
    done
  done
fi
