#!/bin/bash
# same-box A/B of ab/libocb_old.so against ab/libocb_new.so on the fused self-play rollout (parity of `new` first)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu 2>&1 | tail -2
cp diverse_conventions_b200/libocb.so /tmp/libocb_keep.so
for rep in 1 2; do for v in old new; do
  cp ab/libocb_$v.so diverse_conventions_b200/libocb.so; touch diverse_conventions_b200/libocb.so
  echo "== $v"
  timeout 300 python tools/rollout_bench.py --mode selfplay --layouts ${LAYOUTS:-simple,random1,unident_s} --worlds 8192 --T 100 --fused 1 2>&1 | cut -c1-230
done; done 2>&1 | tee gpurun_out/ab_fused.txt
cp /tmp/libocb_keep.so diverse_conventions_b200/libocb.so
python tools/fused_trace.py 2>&1 | tail -1 | cut -c1-1500
