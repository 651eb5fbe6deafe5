#!/bin/bash
# round 2, call A: reference-trainer drop-in tests, fused-rollout step trace, MMA issue-rate probe
mkdir -p gpurun_out
python -m pytest tests/test_gpu_reference_trainers.py -x -q > gpurun_out/r2a_reftests.log 2>&1; echo "reftests rc=$?"
tail -15 gpurun_out/r2a_reftests.log
for L in simple random1 unident_s; do
  python tools/fused_trace.py --layout $L --worlds 8192 --T 40 --u0 8 --steps 8 > gpurun_out/r2a_trace_$L.jsonl 2>&1
  tail -1 gpurun_out/r2a_trace_$L.jsonl
done
./tools/probes/umma_probe2_probe > gpurun_out/r2a_umma_probe2.txt 2>&1; cat gpurun_out/r2a_umma_probe2.txt
python tools/rollout_bench.py --mode selfplay --layouts simple,random1 --worlds 8192 --T 100 > gpurun_out/r2a_selfplay.jsonl 2>&1; cat gpurun_out/r2a_selfplay.jsonl
