#!/bin/bash
# quick GPU iteration: parity tests, then the single-step env kernel sweep and the cross-play matrix
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/it_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/it_pytest.log
for g in 1 2 4 8; do python tools/step_single.py --lanes $g; done
for g in 1 2 4 8; do python tools/step_single.py --lanes $g --worlds 32768; done
for g in 1 2 4 8; do python tools/step_single.py --lanes $g --worlds 8192 --layout simple; done
for g in 2 4 8; do python tools/step_single.py --lanes $g --worlds 1024 --layout simple; done
timeout 600 python tools/rollout_bench.py --mode crossplay --policies 16 --worlds-per-pair 1024 2>&1 | tail -1
python bench.py --no-config4 --no-config5 2>&1 | tail -1 | cut -c1-900
