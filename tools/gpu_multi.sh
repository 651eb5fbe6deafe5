#!/bin/bash
# N-GPU pass (N = $1): headline bench, self-play rollout and the cross-play pair matrix under torchrun.
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 1000 --warmup 20 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench exit $?"; cat gpurun_out/bench_${N}gpu.json
$TR tools/rollout_bench.py --mode selfplay --layouts simple,random1 --worlds 8192 --T 100 > gpurun_out/selfplay_${N}gpu.jsonl 2>&1; tail -2 gpurun_out/selfplay_${N}gpu.jsonl
$TR tools/rollout_bench.py --mode crossplay --policies 16 --worlds-per-pair 1024 > gpurun_out/crossplay_${N}gpu.jsonl 2>&1; tail -1 gpurun_out/crossplay_${N}gpu.jsonl
python bench.py --impl reference --gpus $N --steps 4 --warmup 3 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
