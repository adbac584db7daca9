#!/bin/bash
# Build libhelm_b200.so in-tree for sm_100a (B200).  Usage: build.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${HELM_B200_OUT:-$HERE/../libhelm_b200.so}"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
     -Xcompiler -fPIC -shared "$@" \
     -o "$OUT" "$HERE/helm_b200.cu" "$HERE/symbolic.cpp" "$HERE/symbolic_capi.cpp"
echo "built $OUT"
