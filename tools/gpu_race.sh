#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_overcooked.py tests/test_gpu_rollout.py tests/test_gpu_mixed.py -x -q > gpurun_out/pytest_env.log 2>&1; echo "env pytest exit $?"; tail -3 gpurun_out/pytest_env.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 --launch-timeout 0 python -m pytest tests/test_gpu_overcooked.py -k "golden or random" -x -q > gpurun_out/sanitizer_race_env.log 2>&1; echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_race_env.log | tail -3
python tools/sweep.py --layouts simple,unident_s --worlds 4096,8192,16384 --passes 60 > gpurun_out/sweep_private.jsonl 2>&1
python tools/step_latency.py > gpurun_out/step_latency.txt 2>&1; tail -6 gpurun_out/step_latency.txt
echo done
