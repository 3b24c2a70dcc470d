#!/bin/bash
# build macro variants of the library on the GPU box and bench each (exploration only)
cd "$(dirname "$0")/.."
PK=iterative-linear-quadratic-regulator_b200
for v in "$@"; do
  name=$(echo "$v" | tr -c 'A-Za-z0-9=\n' '_')
  out=/tmp/libilqr_$name.so
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Iinclude -I$PK/csrc $v -o $out $PK/csrc/ilqr_b200.cu || exit 1
  echo "== $v"
  ILQR_B200_LIB=$out python bench.py --no-cpu-baseline --steps 10 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k:(round(v['avg_ms'],3) if isinstance(v,dict) else round(v,3)) for k,v in d['kernels'].items()})"
done
