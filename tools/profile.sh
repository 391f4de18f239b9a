#!/usr/bin/env bash
# Run on the GPU box (under gpurun): plain bench, ncu launch list, full captures of the dominant kernels.
# Cooperative + cluster launches are rejected under Nsight Compute, so the backward kernel is launched with the cluster
# attribute only (PAULE_NO_COOP_CLUSTER=1; its 96 CTAs are co-resident on the otherwise idle GPU either way).
set -u
mkdir -p gpurun_out
export PAULE_NO_COOP_CLUSTER=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
tail -c 600 gpurun_out/prof_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches.csv)"
for spec in "bwd:tc_lstm_bwd2_kernel:3" "fwd:tc_lstm_fwd2_kernel:3" "gemm:tc_gemm_img_kernel:4" "elem:adam_clamp_kernel|smooth_terms_kernel|word_loss_kernel:3"; do
  name=${spec%%:*}; rest=${spec#*:}; rx=${rest%%:*}; skip=${rest##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $([ "$name" = elem ] && echo 3 || echo 1) \
      -o gpurun_out/prof_$name -f $CMD > gpurun_out/ncu_$name.log 2>&1
  echo "$name capture rc=$?"; tail -2 gpurun_out/ncu_$name.log
done
ls -la gpurun_out/*.ncu-rep
