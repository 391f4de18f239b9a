#!/usr/bin/env bash
# Build libpaule_b200.so (C ABI, include/paule_b200.h) for sm_100a, in-tree.
# Every .cu is compiled to its own object in parallel (ptxas -v output kept per file in paule_b200/lib/ptxas.log), then
# linked into one shared library.
set -euo pipefail
cd "$(dirname "$0")"
OUT=paule_b200/lib
OBJ=build/obj
mkdir -p "$OUT" "$OBJ"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-fvisibility=hidden -Xptxas -v ${NVCC_EXTRA:-}"
# a failed build must never leave a stale library behind
rm -f "$OUT/libpaule_b200.so.tmp" "$OBJ"/*.o "$OBJ"/*.log
pids=()
for src in paule_b200/csrc/*.cu; do
  base=$(basename "$src" .cu)
  ( nvcc $FLAGS -c "$src" -o "$OBJ/$base.o" 2> "$OBJ/$base.log" ) &
  pids+=($!)
done
fail=0
for p in "${pids[@]}"; do wait "$p" || fail=1; done
cat "$OBJ"/*.log > "$OUT/ptxas.log"
if [ "$fail" -ne 0 ] || ! nvcc -gencode arch=compute_100a,code=sm_100a --shared -o "$OUT/libpaule_b200.so.tmp" "$OBJ"/*.o 2>> "$OUT/ptxas.log"; then
  grep -E "error" -A3 "$OUT/ptxas.log" | head -40
  rm -f "$OUT/libpaule_b200.so" "$OUT/libpaule_b200.so.tmp"
  echo "BUILD FAILED"
  exit 1
fi
mv "$OUT/libpaule_b200.so.tmp" "$OUT/libpaule_b200.so"
grep -E "warning" "$OUT/ptxas.log" | grep -v "ptxas info" | sort -u | head -5 || true
echo "built $OUT/libpaule_b200.so"
