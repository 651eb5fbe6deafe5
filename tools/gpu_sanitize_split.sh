#!/bin/bash
# compute-sanitizer over the role-split env kernel (named-barrier ring between the transition and the encoder warps) and the
# narrow-tile launches: memcheck, racecheck (shared-memory hazards), synccheck (barrier misuse)
set -u
mkdir -p gpurun_out
run() { tool=$1; name=$2; to=$3; shift 3; timeout $to compute-sanitizer --tool $tool --error-exitcode 99 --launch-timeout 0 python -m pytest "$@" -x -q > gpurun_out/r2g_sanitizer_${tool}_$name.log 2>&1; echo "$tool $name exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2g_sanitizer_${tool}_$name.log | tail -3; }
run memcheck split 600 tests/test_gpu_overcooked.py -k "(random_rollout and 16) or narrow_tiles"
run racecheck split 600 tests/test_gpu_overcooked.py -k "(random_rollout and 16-1 and (simple or unident_s or random0)) or (narrow_tiles and 16)"
run synccheck split 600 tests/test_gpu_overcooked.py -k "(random_rollout and 16-1 and (simple or unident_s)) or (narrow_tiles and 16)"
for f in gpurun_out/r2g_sanitizer_*.log; do tail -c 1500 $f > $f.tail; mv $f.tail $f; done
echo done
