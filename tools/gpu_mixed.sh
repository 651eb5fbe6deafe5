#!/bin/bash
# Mixed-play collection: parity first, timings, then the full GPU suite, smoke and the headline bench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mixed.py -x -q > gpurun_out/pytest_mixed.log 2>&1; echo "mixed pytest exit $?"; tail -30 gpurun_out/pytest_mixed.log
timeout 600 python tools/rollout_bench.py --mode mixed --layouts simple,random1 2>&1 | tee gpurun_out/mixed_bench.jsonl
timeout 600 python tools/rollout_bench.py --mode mixed --layouts simple --graph 0 2>&1 | tee -a gpurun_out/mixed_bench.jsonl
timeout 600 python tools/rollout_bench.py --mode mixed --layouts simple --hidden 512 --replicas 10 2>&1 | tee -a gpurun_out/mixed_bench.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json
echo done
