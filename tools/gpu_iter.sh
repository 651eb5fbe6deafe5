#!/bin/bash
set -u
mkdir -p gpurun_out
for s in 1 2; do
  echo "OCB_FUSED_SLOTS=$s"
  OCB_FUSED_SLOTS=$s timeout 600 python tools/rollout_bench.py --mode crossplay --policies 16 --worlds-per-pair 1024 2>&1 | tail -1 | cut -c1-330
  OCB_FUSED_SLOTS=$s timeout 600 python tools/rollout_bench.py --mode selfplay --layouts simple,random1 --worlds 32768 --T 50 2>&1 | tail -2
done
timeout 600 python tools/rollout_bench.py --mode selfplay --layouts simple,random1 --worlds 8192 --T 100 2>&1 | tail -2
