#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ppo.py -x -q > gpurun_out/pytest_ppo.log 2>&1; echo "ppo pytest exit $?"; tail -5 gpurun_out/pytest_ppo.log
timeout 600 python tools/ppo_bench.py --layouts simple,random3 2>&1 | tee gpurun_out/ppo_bench_realign.jsonl
OCB_GATHER_VEC16=1 timeout 600 python tools/ppo_bench.py --layouts simple,random3 2>&1 | tee gpurun_out/ppo_bench_vec16.jsonl
echo done
