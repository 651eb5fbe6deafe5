#!/bin/bash
# Evidence pass: PPO minibatch timings, role-level stall breakdown of the policy kernel, launch list of the mixed-play collection
set -u
mkdir -p gpurun_out
timeout 600 python tools/ppo_bench.py 2>&1 | tee gpurun_out/ppo_bench.jsonl
timeout 300 python tools/policy_roles.py --rows 32768 > gpurun_out/roles_32768.txt 2>&1; tail -40 gpurun_out/roles_32768.txt
timeout 300 python tools/policy_roles.py --rows 262144 > gpurun_out/roles_262144.txt 2>&1; tail -40 gpurun_out/roles_262144.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4100 -c 60 --csv --log-file gpurun_out/launches_mixed.csv \
  python tools/rollout_bench.py --mode mixed --layouts simple --graph 0 --iters 1 > gpurun_out/ncu_mixed.log 2>&1; tail -2 gpurun_out/ncu_mixed.log
echo done
