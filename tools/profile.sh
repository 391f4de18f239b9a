#!/usr/bin/env bash
# Run on the GPU box (under gpurun): plain bench, ncu launch list, one full capture of the dominant kernels.
set -u
mkdir -p gpurun_out
export PAULE_NO_COOP_CLUSTER=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
tail -c 600 gpurun_out/prof_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches.csv)"
ncu --set full --clock-control none --import-source on -k regex:tc_lstm_bwd_kernel -s 3 -c 1 -o gpurun_out/prof_bwd -f $CMD > gpurun_out/ncu_bwd.log 2>&1
echo "bwd capture rc=$?"; tail -3 gpurun_out/ncu_bwd.log
ls -la gpurun_out/*.ncu-rep
