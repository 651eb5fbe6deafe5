#!/bin/bash
# Re-validation pass: full GPU parity suite, smoke, headline bench (+ reference arm), fused-rollout timing and ncu evidence.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 4 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cat gpurun_out/bench_ref.json
for f in 0 1; do
  timeout 300 python tools/rollout_bench.py --mode selfplay --worlds 8192 --T 100 --fused $f 2>&1 | tee -a gpurun_out/rollout_fused_cmp.jsonl
done
timeout 300 python tools/rollout_bench.py --mode selfplay --layouts simple --worlds 16384 --T 100 --fused 1 2>&1 | tee -a gpurun_out/rollout_fused_cmp.jsonl
timeout 300 python tools/rollout_bench.py --mode selfplay --layouts simple --worlds 9472 --T 100 --fused 1 2>&1 | tee -a gpurun_out/rollout_fused_cmp.jsonl
timeout 300 python tools/rollout_bench.py --mode selfplay --worlds 8192 --T 100 --hidden 512 2>&1 | tee gpurun_out/rollout_selfplay_h512.jsonl
timeout 300 python tools/rollout_bench.py --mode crossplay --policies 16 --worlds-per-pair 128 2>&1 | tee gpurun_out/rollout_crossplay.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_fused -c 1 -o gpurun_out/rollout_fused_full -f \
  python tools/rollout_bench.py --mode selfplay --layouts simple --worlds 8192 --T 100 --fused 1 --iters 1 --graph 0 > gpurun_out/ncu_rollout_fused.log 2>&1; tail -3 gpurun_out/ncu_rollout_fused.log
echo done
