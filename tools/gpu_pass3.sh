#!/bin/bash
# PPO minibatch path: parity first, then the full GPU suite, timings and ncu evidence
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ppo.py -x -q > gpurun_out/pytest_ppo.log 2>&1; echo "ppo pytest exit $?"; tail -25 gpurun_out/pytest_ppo.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/ppo_bench.py 2>&1 | tee gpurun_out/ppo_bench.jsonl
timeout 600 python tools/ppo_bench.py --layouts simple --hidden 512 2>&1 | tee -a gpurun_out/ppo_bench.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:minibatch_gather -s 3 -c 2 -o gpurun_out/gather_full -f \
  python tools/ppo_bench.py --layouts simple --iters 2 > gpurun_out/ncu_gather.log 2>&1; tail -2 gpurun_out/ncu_gather.log
echo done
