#!/bin/bash
# second compute-sanitizer pass: tensor-core policy kernels, rollouts, PPO path (memcheck), env kernel (racecheck)
set -u
mkdir -p gpurun_out
CS="compute-sanitizer --error-exitcode 99 --launch-timeout 0"
run() { name=$1; tool=$2; shift 2; timeout $1 $CS --tool $tool python -m pytest "${@:2}" -x -q > gpurun_out/sanitizer_$name.log 2>&1; echo "$name ($tool) exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|hazard" gpurun_out/sanitizer_$name.log | tail -4; }
run policy memcheck 420 tests/test_gpu_policy.py
run ppo memcheck 420 tests/test_gpu_ppo.py
run rollout memcheck 600 tests/test_gpu_rollout.py
run policy512 memcheck 420 tests/test_gpu_policy512.py
run race_env racecheck 600 tests/test_gpu_overcooked.py -k "golden or random"
run race_small racecheck 420 tests/test_gpu_balance.py tests/test_gpu_returns.py "tests/test_gpu_ppo.py::test_gather_large_ragged_against_torch_indexing"
echo done
