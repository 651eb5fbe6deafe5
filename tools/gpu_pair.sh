#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_policy.py tests/test_gpu_rollout.py tests/test_gpu_mixed.py tests/test_gpu_ppo.py -x -q > gpurun_out/pytest_pair.log 2>&1; echo "pair pytest exit $?"; tail -5 gpurun_out/pytest_pair.log
timeout 300 python tools/policy_bench.py 2>&1 | tee gpurun_out/policy_bench.jsonl | tail -8
timeout 300 python tools/rollout_bench.py --mode selfplay --layouts simple,random1,unident_s --worlds 8192 --T 100 2>&1 | tee gpurun_out/selfplay_pair.jsonl
timeout 300 python tools/rollout_bench.py --mode selfplay --layouts simple --worlds 8192 --T 100 --fused 0 2>&1 | tee -a gpurun_out/selfplay_pair.jsonl
timeout 300 python tools/policy_roles.py --rows 262144 > gpurun_out/roles_262144.txt 2>&1; tail -6 gpurun_out/roles_262144.txt
echo done
