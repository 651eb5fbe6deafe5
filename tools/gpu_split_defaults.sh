#!/bin/bash
# which launch shape wins where: role-split kernel (16; GE = 2 / 4) against the one-warp kernel (1, 2, 4), per layout and world count
mkdir -p gpurun_out
for layout in simple random1 unident_s; do
  W=1024,2048,4096,8192,12288,16384,32768
  for ge in 2 4; do OCB_SPLIT_GE=$ge timeout 600 python tools/sweep.py --layouts $layout --worlds $W --lanes 16 --quick --tma 1 --passes 10; done
  timeout 600 python tools/sweep.py --layouts $layout --worlds $W --lanes 1,2,4 --quick --tma 1 --passes 10
done 2>&1 | tee gpurun_out/split_defaults.jsonl | python -c "
import sys, json
rows=[json.loads(l) for l in sys.stdin if l.startswith('{')]
best={}
for r in rows:
    k=(r['layout'],r['N']); best.setdefault(k,[]).append((r['ms'], r['G'], r['split_ge'], r['frac']))
for k in best:
    print(k, sorted(best[k]))
"
