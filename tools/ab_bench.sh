#!/bin/bash
# same-box A/B of builds of libocb.so (ab/libocb_<name>.so) on the headline kernel and on the single-step launch
set -u
cp diverse_conventions_b200/libocb.so /tmp/libocb_keep.so
for rep in 1 2; do
for f in ab/libocb_*.so; do  # build the variants into ab/ first (git-ignored; they travel with gpurun)
  cp $f diverse_conventions_b200/libocb.so; touch diverse_conventions_b200/libocb.so
  echo "== $f"; python bench.py --no-config4 --no-config5 --no-cpu-baseline --steps 200 --warmup 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('K=100 launch_ms', round(d['roofline']['launch_ms'],4), 'frac', round(d['roofline']['frac'],3))"
  python tools/step_single.py --lanes 2 2>&1 | tail -1 | cut -c1-90
done; done
cp /tmp/libocb_keep.so diverse_conventions_b200/libocb.so
