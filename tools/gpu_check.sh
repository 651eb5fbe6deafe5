#!/bin/bash
# One GPU-box pass: parity tests, smoke, headline bench, policy / rollout benches, ncu evidence.
# Everything lands in gpurun_out/ (scratch); summaries are copied into profiles/ by hand.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json
python tools/policy_bench.py --layouts simple,unident_s,random1,random0,random3 --mode fused > gpurun_out/policy_bench.jsonl 2>&1; cat gpurun_out/policy_bench.jsonl
python tools/rollout_bench.py --mode selfplay --worlds 8192 --T 100 > gpurun_out/rollout_selfplay.jsonl 2>&1; cat gpurun_out/rollout_selfplay.jsonl
python tools/rollout_bench.py --mode crossplay --policies 16 --worlds-per-pair 128 > gpurun_out/rollout_crossplay.jsonl 2>&1; cat gpurun_out/rollout_crossplay.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_rollout.csv \
  python tools/rollout_bench.py --mode selfplay --layouts simple --worlds 8192 --T 50 --iters 1 --graph 0 > gpurun_out/launches_rollout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:policy -s 4 -c 2 -o gpurun_out/policy_full -f \
  python tools/policy_bench.py --layouts simple --mode fused --iters 5 > gpurun_out/ncu_policy.log 2>&1
python tools/policy512_bench.py --layouts simple,unident_s,random1,random0,random3 > gpurun_out/policy512_bench.jsonl 2>&1; cat gpurun_out/policy512_bench.jsonl
python tools/rollout_bench.py --mode selfplay --worlds 8192 --T 100 --hidden 512 > gpurun_out/rollout_selfplay_h512.jsonl 2>&1; cat gpurun_out/rollout_selfplay_h512.jsonl
ncu --set full --clock-control none --import-source on -k regex:'conv512|gemm512' -s 6 -c 3 -o gpurun_out/policy512_full -f \
  python tools/policy512_bench.py --layouts simple --iters 5 > gpurun_out/ncu_policy512.log 2>&1
echo done
