#!/bin/bash
# ncu --set full of one Balance-Beam launch (65,536 worlds, K = 100): per-source-line instruction / stall shares + stall reasons
set -u
mkdir -p gpurun_out /tmp/cub && (cd /tmp/cub && cuobjdump -xelf all $OLDPWD/diverse_conventions_b200/libocb.so > /dev/null 2>&1)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bb_kernel -s 8 -c 1 -o gpurun_out/bb_full -f \
  python tools/bb_bench.py --worlds 65536 --iters 5 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/bb_full.ncu-rep gpurun_out/bb_full_summary.json > /dev/null
python tools/ncu_lines_by_source.py gpurun_out/bb_full.ncu-rep /tmp/cub/bb_kernels.sm_100a.cubin 9bb_kernelE 70 > gpurun_out/bb_lines.txt 2>&1
ncu -i gpurun_out/bb_full.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; units=rows[1]; vals=rows[2]
for h,u,v in zip(hdr,units,vals):
    if ('issue_stalled' in h and 'per_warp_active.pct' in h) or h in ('smsp__inst_executed.sum','gpu__time_duration.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes_write.sum','dram__bytes_read.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size'): print(h,u,v)
" > gpurun_out/bb_stalls.txt
rm -f gpurun_out/bb_full.ncu-rep
head -75 gpurun_out/bb_lines.txt; cat gpurun_out/bb_stalls.txt
