#!/bin/bash
# third compute-sanitizer pass: initcheck (reads of uninitialised device memory) and synccheck (barrier misuse)
set -u
mkdir -p gpurun_out
CS="compute-sanitizer --error-exitcode 99 --launch-timeout 0"
run() { name=$1; tool=$2; shift 2; timeout $1 $CS --tool $tool python -m pytest "${@:2}" -x -q > gpurun_out/sanitizer_$name.log 2>&1; echo "$name ($tool) exit $?"; grep -E "ERROR SUMMARY|passed|failed|Uninitialized|Barrier error" gpurun_out/sanitizer_$name.log | sort | uniq -c | tail -5; }
run init_env initcheck 600 tests/test_gpu_overcooked.py -k "golden or random or inject"
run init_small initcheck 420 tests/test_gpu_balance.py tests/test_gpu_returns.py tests/test_gpu_ppo.py tests/test_gpu_mixed.py
run init_rollout initcheck 600 tests/test_gpu_rollout.py tests/test_gpu_policy.py
run sync_env synccheck 420 tests/test_gpu_overcooked.py -k "golden or random"
run sync_policy synccheck 600 tests/test_gpu_policy.py tests/test_gpu_rollout.py tests/test_gpu_policy512.py
echo done
