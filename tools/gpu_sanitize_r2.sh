#!/bin/bash
# round 2: compute-sanitizer memcheck / racecheck over the tests of the kernels that changed this round
# (single-step env instance with the bulk-copy rebuild, single-network policy mode, fused rollout with one / two tiles, cross-play)
set -u
mkdir -p gpurun_out
run() { tool=$1; name=$2; to=$3; shift 3; timeout $to compute-sanitizer --tool $tool --error-exitcode 99 --launch-timeout 0 python -m pytest "$@" -x -q > gpurun_out/r2_sanitizer_${tool}_$name.log 2>&1; echo "$tool $name exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2_sanitizer_${tool}_$name.log | tail -3; }
run memcheck env_single 500 tests/test_gpu_overcooked.py -k "single_steps or step_api or observe or state_injection"
run memcheck policy_single 400 tests/test_gpu_policy.py -k "single_network or golden or ragged"
run memcheck fused 700 tests/test_gpu_rollout.py -k "bit_identical or crossplay_slices"
run racecheck env_single 500 tests/test_gpu_overcooked.py -k "single_steps"
for f in gpurun_out/r2_sanitizer_*.log; do tail -c 2000 $f > $f.tail; mv $f.tail $f; done
echo done
