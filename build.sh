#!/usr/bin/env bash
# Build libpaule_b200.so (C ABI, include/paule_b200.h) for sm_100a, in-tree.
set -euo pipefail
cd "$(dirname "$0")"
OUT=paule_b200/lib
mkdir -p "$OUT"
SRCS=$(ls paule_b200/csrc/*.cu)
# build into a temporary name: a failed build must never leave a stale library behind
rm -f "$OUT/libpaule_b200.so.tmp"
if ! nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
     -Xcompiler -fPIC,-Wall,-fvisibility=hidden --shared \
     -Xptxas -v ${NVCC_EXTRA:-} \
     -o "$OUT/libpaule_b200.so.tmp" $SRCS 2> "$OUT/ptxas.log"; then
  grep -E "error" -A3 "$OUT/ptxas.log" | head -40
  rm -f "$OUT/libpaule_b200.so" "$OUT/libpaule_b200.so.tmp"
  echo "BUILD FAILED"
  exit 1
fi
mv "$OUT/libpaule_b200.so.tmp" "$OUT/libpaule_b200.so"
grep -E "warning" "$OUT/ptxas.log" | grep -v "ptxas info" | head -5 || true
echo "built $OUT/libpaule_b200.so"
