#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_policy.py tests/test_gpu_rollout.py tests/test_gpu_ppo.py tests/test_gpu_mixed.py -x -q 2>&1 | tail -3
for v in 0 1; do
  echo "OCB_D3_ON_D2=$v"
  OCB_D3_ON_D2=$v timeout 300 python tools/policy_bench.py --layouts simple,random1,unident_s --mode fused 2>&1 | tail -3
  OCB_D3_ON_D2=$v timeout 300 python tools/policy_bench.py --layouts simple --rows 262144 --mode fused 2>&1 | tail -1
  OCB_D3_ON_D2=$v timeout 600 python tools/rollout_bench.py --mode crossplay --policies 16 --worlds-per-pair 1024 2>&1 | tail -1 | cut -c1-200
  OCB_D3_ON_D2=$v timeout 600 python tools/rollout_bench.py --mode selfplay --layouts simple,random1 --worlds 32768 --T 50 2>&1 | tail -2 | cut -c1-260
done
