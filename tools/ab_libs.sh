#!/usr/bin/env bash
# A/B of builds of the library (PAULE_B200_LIB): recurrent kernels alone at several batch sizes + one planner step
set -u
OUT=gpurun_out/${1:-ab_libs}.txt
: > $OUT
for lib in ${LIBS:-libpaule_b200_nohint libpaule_b200 libpaule_b200_hint20k}; do
  [ -f paule_b200/lib/$lib.so ] || continue
  echo "== $lib" | tee -a $OUT
  PAULE_B200_LIB=$PWD/paule_b200/lib/$lib.so python tools/rnn_time.py ${SIZES:-1 64 128 256 384 1024} 2>&1 | tee -a $OUT
  PAULE_B200_LIB=$PWD/paule_b200/lib/$lib.so python tools/fwd_time.py ${STEP_SIZES:-1 64 256} 2>&1 | tee -a $OUT
done
