#!/bin/bash
# fused rollout kernel: parity first, then timing against the per-step launches
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rollout.py -x -q > gpurun_out/pytest_fused.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/pytest_fused.log
for f in 0 1; do
  timeout 300 python tools/rollout_bench.py --mode selfplay --worlds 8192 --T 100 --fused $f 2>&1 | tee -a gpurun_out/rollout_fused_cmp.jsonl
done
timeout 300 python tools/rollout_bench.py --mode selfplay --layouts simple --worlds 16384 --T 100 --fused 1 2>&1 | tee -a gpurun_out/rollout_fused_cmp.jsonl
timeout 300 python tools/rollout_bench.py --mode selfplay --layouts simple --worlds 9472 --T 100 --fused 1 2>&1 | tee -a gpurun_out/rollout_fused_cmp.jsonl
echo done
