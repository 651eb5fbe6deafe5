#!/bin/bash
# role-split env kernel: parity of the new variant, then the timing sweep against the one-warp kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_overcooked.py tests/test_gpu_random_layouts.py -x -q -m gpu -k "16 or split" 2>&1 | tail -3
for ge in 4 2; do
  OCB_SPLIT_GE=$ge timeout 300 python tools/sweep.py --layouts simple --worlds ${WORLDS:-8192,16384} --lanes 16 --quick --tma 1 --T ${T:-100}
done 2>&1 | tee gpurun_out/split_sweep.jsonl
timeout 300 python tools/sweep.py --layouts simple --worlds ${WORLDS:-8192,16384} --lanes 1,4 --quick --tma 1 --T ${T:-100} 2>&1 | tee -a gpurun_out/split_sweep.jsonl
