#!/usr/bin/env bash
# Per-step time of a fixed multi-slot layout of the forward recurrent kernel as a function of the number of co-resident word
# groups (1..6): flat = bound inside the CTA (loader visits, epilogue passes, exchange latency), growing = bound by a shared
# resource (L2 bandwidth, polling pressure).  Round-2 result (profiles/r2_nonblocking_multislot_ab.txt): flat, 5.6 us per step
# for 64 words per CTA row with 1..6 groups.
set -u
OUT=gpurun_out/${1:-ab_groups}.txt
: > $OUT
for L in ${LAYOUTS:-222 311 212}; do
  echo "== layout $L" | tee -a $OUT
  PAULE_FWD_LAYOUT=$L python tools/rnn_time.py 64 128 192 256 320 384 2>&1 | tee -a $OUT
done
