#!/usr/bin/env bash
# Run on the GPU box (under gpurun): plain bench, ncu launch list of one inner step, full captures of the dominant kernels.
# Nsight Compute serialises kernels (kernel replay); the layer wavefront launches its kernels in dependency order, so every
# waiter finds its counters complete and the serialised run is correct (only slower).
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --quick --min-timed-s 0"
$CMD > gpurun_out/${TAG}_prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_prof_plain.log; exit 1; }
tail -c 400 gpurun_out/${TAG}_prof_plain.log; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/${TAG}_launches.csv)"
for spec in "bwd:tc_lstm_bwd2_kernel:4" "fwd:tc_lstm_fwd2_kernel:3" "gemm:tc_gemm_img_kernel:30" "gemm2:tc_gemm_img2_kernel:4" "elem:adam_clamp_kernel|smooth_terms_kernel|word_loss_kernel:3"; do
  name=${spec%%:*}; rest=${spec#*:}; rx=${rest%%:*}; skip=${rest##*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $([ "$name" = elem ] && echo 3 || echo 1) \
      -o gpurun_out/${TAG}_prof_$name -f $CMD > gpurun_out/${TAG}_ncu_$name.log 2>&1
  echo "$name capture rc=$?"; tail -2 gpurun_out/${TAG}_ncu_$name.log
done
# the gate GEMM of embedder layer 1 (K = 720, N = 2880) as a batch GEMM at 256 words: the instance bench.py's kernel_rooflines times
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:tc_gemm_img2_kernel" -s 1 -c 1 \
    -o gpurun_out/${TAG}_prof_gategemm -f python tools/gemm_time.py 256 > gpurun_out/${TAG}_ncu_gategemm.log 2>&1
echo "gate GEMM capture rc=$?"
ls -la gpurun_out/${TAG}_*.ncu-rep
