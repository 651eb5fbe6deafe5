#!/bin/bash
# One GPU-box pass: parity tests, smoke, headline bench (both arms), policy / rollout / env benches.
# Everything lands in gpurun_out/ (scratch); summaries are copied into profiles/ by hand.  ncu evidence: tools/gpu_r2_profiles.sh;
# compute-sanitizer: tools/gpu_sanitize_r2.sh; same-box A/B of two builds: tools/ab_bench.sh.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-600 gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference arm exit $?"
python tools/policy_bench.py --layouts simple,unident_s,random1,random0,random3 --mode fused > gpurun_out/policy_bench.jsonl 2>&1; cat gpurun_out/policy_bench.jsonl
python tools/policy_bench.py --layouts random1 --rows 524288 --mode act >> gpurun_out/policy_bench.jsonl 2>&1; tail -1 gpurun_out/policy_bench.jsonl
python tools/rollout_bench.py --mode selfplay --worlds 8192 --T 100 > gpurun_out/rollout_selfplay.jsonl 2>&1; cat gpurun_out/rollout_selfplay.jsonl
python tools/rollout_bench.py --mode selfplay --layouts simple,random1 --worlds 32768 --T 50 > gpurun_out/rollout_selfplay_32k.jsonl 2>&1; cat gpurun_out/rollout_selfplay_32k.jsonl
python tools/rollout_bench.py --mode crossplay --policies 16 --worlds-per-pair 1024 > gpurun_out/rollout_crossplay.jsonl 2>&1; cut -c1-300 gpurun_out/rollout_crossplay.jsonl
for w in 1024 8192 32768 262144; do python tools/step_single.py --worlds $w; done > gpurun_out/step_single.jsonl 2>&1; cat gpurun_out/step_single.jsonl
python tools/bb_bench.py > gpurun_out/bb_bench.jsonl 2>&1; cat gpurun_out/bb_bench.jsonl
python tools/policy512_bench.py --layouts simple,unident_s,random1,random0,random3 > gpurun_out/policy512_bench.jsonl 2>&1; cat gpurun_out/policy512_bench.jsonl
python tools/rollout_bench.py --mode mixed > gpurun_out/mixed_bench.jsonl 2>&1; cat gpurun_out/mixed_bench.jsonl
python tools/ppo_bench.py > gpurun_out/ppo_bench.jsonl 2>&1; tail -5 gpurun_out/ppo_bench.jsonl
echo done
