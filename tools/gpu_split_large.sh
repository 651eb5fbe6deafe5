#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
for layout in simple random1 random3; do
  W=16384,32768,65536,262144
  OCB_SPLIT_GE=2 timeout 600 python tools/sweep.py --layouts $layout --worlds $W --lanes 16 --quick --tma 1 --passes 10
  timeout 600 python tools/sweep.py --layouts $layout --worlds $W --lanes 1,2 --quick --tma 1 --passes 10
  OCB_TILE_WORLDS=32 timeout 600 python tools/sweep.py --layouts $layout --worlds 16384 --lanes 1,2 --quick --tma 1 --passes 10 | sed 's/"layout"/"tile32": 1, "layout"/'
done 2>&1 | tee gpurun_out/split_large.jsonl | cut -c1-200
