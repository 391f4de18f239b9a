#!/usr/bin/env bash
# ncu --set full captures (with source counters) of the multi-quarter layouts of the recurrent kernels and of the gate GEMM
set -u
mkdir -p gpurun_out
TAG=${1:-r2b}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_lstm_fwd2_kernel -s 1 -c 1 -o gpurun_out/${TAG}_prof_fwd384 -f python tools/rnn_time.py 384 > gpurun_out/${TAG}_ncu_fwd384.log 2>&1; echo "fwd384 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_lstm_bwd2_kernel -s 1 -c 1 -o gpurun_out/${TAG}_prof_bwd320 -f python tools/rnn_time.py 320 > gpurun_out/${TAG}_ncu_bwd320.log 2>&1; echo "bwd320 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_img_kernel -s 1 -c 1 -o gpurun_out/${TAG}_prof_gategemm -f python tools/gemm_time.py 64 > gpurun_out/${TAG}_ncu_gategemm.log 2>&1; echo "gategemm rc=$?"
ls -la gpurun_out/${TAG}_*.ncu-rep
