#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/it_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/it_pytest.log
timeout 300 python tools/policy_bench.py --layouts simple,random1 --rows 32768 --mode act 2>&1 | tail -2
timeout 300 python tools/policy_bench.py --layouts random1 --rows 524288 --mode act 2>&1 | tail -1
timeout 300 python tools/policy_bench.py --layouts random1 --rows 524288 --mode value 2>&1 | tail -1
timeout 600 python tools/rollout_bench.py --mode crossplay --policies 16 --worlds-per-pair 1024 2>&1 | tail -1 | cut -c1-330
bash tools/ab_bench.sh
