#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_overcooked.py -x -q -m gpu -k "narrow or full_size or random_rollout" 2>&1 | tail -3
for layout in ${LAYOUTS:-simple random3}; do
for tw in 8 16 24 32 auto; do
  [ $tw = auto ] && unset OCB_TILE_WORLDS || export OCB_TILE_WORLDS=$tw
  for ge in 2 4; do
    OCB_SPLIT_GE=$ge timeout 300 python tools/sweep.py --layouts $layout --worlds ${WORLDS:-8192,16384,32768} --lanes 16 --quick --tma 1 --T ${T:-100} | sed "s/\"layout\"/\"tile\": \"$tw\", \"layout\"/"
  done
done
unset OCB_TILE_WORLDS
timeout 300 python tools/sweep.py --layouts $layout --worlds ${WORLDS:-8192,16384,32768} --lanes 1,2 --quick --tma 1 --T ${T:-100}
done 2>&1 | tee gpurun_out/tile_sweep.jsonl | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l); print(r.get('tile','-'), r['layout'], r['N'], 'G', r['G'], 'ge', r['split_ge'], r['ms'], r['frac'])
    else: print(l.strip()[:150])
"
