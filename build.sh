#!/usr/bin/env bash
# Build libpaule_b200.so (C ABI, include/paule_b200.h) for sm_100a, in-tree.
set -euo pipefail
cd "$(dirname "$0")"
OUT=paule_b200/lib
mkdir -p "$OUT"
SRCS=$(ls paule_b200/csrc/*.cu)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
     -Xcompiler -fPIC,-Wall,-fvisibility=hidden --shared \
     -Xptxas -v ${NVCC_EXTRA:-} \
     -o "$OUT/libpaule_b200.so" $SRCS 2> "$OUT/ptxas.log" || { cat "$OUT/ptxas.log"; exit 1; }
grep -E "error|warning" "$OUT/ptxas.log" | grep -v "ptxas info" || true
echo "built $OUT/libpaule_b200.so"
