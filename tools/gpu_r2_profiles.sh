#!/bin/bash
# round 2 ncu evidence: launch list of the bench command, --set full captures of the three dominant kernels
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv \
  python bench.py --steps 10 --warmup 3 --e2e-steps 10 --no-cpu-baseline --no-config5 --policy-T 20 > gpurun_out/r2_ncu_bench.log 2>&1; tail -c 300 gpurun_out/r2_ncu_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oc_rollout -s 3 -c 2 -o gpurun_out/r2_oc_rollout_full -f \
  python bench.py --steps 6 --warmup 3 --e2e-steps 10 --no-cpu-baseline --no-config4 --no-config5 > gpurun_out/r2_ncu_full_oc.log 2>&1; tail -c 200 gpurun_out/r2_ncu_full_oc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_fused -s 1 -c 1 -o gpurun_out/r2_rollout_fused_full -f \
  python tools/rollout_bench.py --mode selfplay --layouts simple --worlds 8192 --T 100 --graph 0 --iters 1 > gpurun_out/r2_ncu_full_fused.log 2>&1; tail -c 300 gpurun_out/r2_ncu_full_fused.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:policy_pair -s 2 -c 1 -o gpurun_out/r2_policy_pair_full -f \
  python tools/policy_bench.py --mode fused --layouts simple --rows 262144 --iters 2 > gpurun_out/r2_ncu_full_pair.log 2>&1; tail -c 300 gpurun_out/r2_ncu_full_pair.log
ls -la gpurun_out/*.ncu-rep
echo done
