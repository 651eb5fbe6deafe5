#!/bin/bash
mkdir -p gpurun_out
for ge in 4 2; do
  OCB_SPLIT_GE=$ge timeout 300 python tools/sweep.py --layouts simple --worlds ${WORLDS:-8192,16384} --lanes 16 --quick --tma 0,1 --T ${T:-100}
done 2>&1 | tee gpurun_out/split_sweep2.jsonl
timeout 300 python tools/sweep.py --layouts simple --worlds ${WORLDS:-8192,16384} --lanes 1 --quick --tma 0,1 --T ${T:-100} 2>&1 | tee -a gpurun_out/split_sweep2.jsonl
