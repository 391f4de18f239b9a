#!/usr/bin/env bash
# debugging aid of the layer wavefront: inner-step time per (batch, forward mode, backward mode, serialised, quarters per BPTT
# CTA, GEMM CTAs per column tile); 0 = the library's own choice
for cfg in ${CFGS:-"64 2 3 0 0 0"}; do
  set -- $cfg
  echo "== B=$1 FWD=$2 BWD=$3 SERIALIZE=$4 BNQ=$5 PAR=$6"
  if [ "$4" = 1 ]; then export PAULE_WAVE_SERIALIZE=1; else unset PAULE_WAVE_SERIALIZE; fi
  [ "$5" != 0 ] && export PAULE_WAVE_BNQ=$5 || unset PAULE_WAVE_BNQ
  [ "$6" != 0 ] && export PAULE_WAVE_PAR=$6 || unset PAULE_WAVE_PAR
  PAULE_WAVEFRONT_FWD=$2 PAULE_WAVEFRONT_BWD=$3 timeout 120 python tools/fwd_time.py $1 2>&1 | tail -2
done
