#!/bin/bash
# BASELINE config 4 at its stated shape (8,192 worlds, T=400, hidden 64 and 512, five layouts) and config 5 on ONE GPU
# (its matrix hash is compared with the sharded runs)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_policy.py -x -q > gpurun_out/pytest_rollout.log 2>&1; echo "rollout pytest exit $?"; tail -4 gpurun_out/pytest_rollout.log
timeout 600 python tools/rollout_bench.py --mode selfplay --worlds 8192 --T 400 --iters 3 2>&1 | tee gpurun_out/config4_T400.jsonl
timeout 900 python tools/rollout_bench.py --mode selfplay --worlds 8192 --T 400 --iters 2 --hidden 512 2>&1 | tee -a gpurun_out/config4_T400.jsonl
timeout 600 python tools/rollout_bench.py --mode crossplay --policies 16 --worlds-per-pair 1024 --iters 2 2>&1 | tee gpurun_out/crossplay_1gpu.jsonl | cut -c1-600
echo done
