#!/bin/bash
# compute-sanitizer memcheck over the GPU parity tests (slow: bounded by timeouts; results under gpurun_out/sanitizer_*.log)
set -u
mkdir -p gpurun_out
CS="compute-sanitizer --tool memcheck --error-exitcode 99 --launch-timeout 0"
run() { name=$1; shift; timeout $1 $CS python -m pytest "${@:2}" -x -q > gpurun_out/sanitizer_$name.log 2>&1; echo "$name exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/sanitizer_$name.log | tail -4; }
run gather 300 tests/test_gpu_ppo.py -k gather
run balance 240 tests/test_gpu_balance.py
run returns 240 tests/test_gpu_returns.py
run mixed 420 tests/test_gpu_mixed.py
run overcooked 600 tests/test_gpu_overcooked.py
echo done
