#!/bin/bash
# refresh the ncu evidence of the headline kernel after the private-column change: launch list of the bench command and
# one --set full capture of oc_rollout_kernel; plus full captures of the new gather / mixed kernels
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2 -c 60 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 10 --warmup 3 --e2e-passes 1 --no-policy-rollout --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oc_rollout -s 3 -c 2 -o gpurun_out/oc_rollout_full -f \
  python bench.py --steps 6 --warmup 3 --e2e-passes 1 --no-policy-rollout --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_realign -s 3 -c 1 -o gpurun_out/gather_realign_full -f \
  python tools/ppo_bench.py --layouts unident_s --iters 2 > gpurun_out/ncu_gather.log 2>&1; tail -1 gpurun_out/ncu_gather.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
echo done
