#!/bin/bash
# ncu --set full of one launch of the role-split env kernel (args: worlds GE TW), key counters + per-line profile
set -u
N=${1:-16384}; export OCB_SPLIT_GE=${2:-4}; export OCB_SPLIT_TW=${3:-1}
TAG=split_${N}_ge${OCB_SPLIT_GE}_tw${OCB_SPLIT_TW}
mkdir -p gpurun_out /tmp/cub && (cd /tmp/cub && cuobjdump -xelf all $OLDPWD/diverse_conventions_b200/libocb.so > /dev/null 2>&1)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oc_rollout_split -s 3 -c 1 -o gpurun_out/$TAG -f \
  python tools/sweep.py --layouts simple --worlds $N --lanes 16 --quick --tma 1 --passes 3 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/$TAG.ncu-rep gpurun_out/${TAG}_summary.json
python tools/ncu_lines_by_source.py gpurun_out/$TAG.ncu-rep /tmp/cub/oc_kernels.sm_100a.cubin oc_rollout_split_kernelILi${OCB_SPLIT_GE}ELi${OCB_SPLIT_TW}E 120 > gpurun_out/${TAG}_lines.txt 2>&1
ncu -i gpurun_out/$TAG.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; units=rows[1]; vals=rows[2]
for h,u,v in zip(hdr,units,vals):
    if ('issue_stalled' in h and 'per_warp_active.pct' in h) or h in ('smsp__inst_executed.sum','gpu__time_duration.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes_write.sum','dram__bytes_read.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__warps_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','lts__t_bytes.sum'): print(h,u,v)
" > gpurun_out/${TAG}_stalls.txt
rm -f gpurun_out/$TAG.ncu-rep
head -125 gpurun_out/${TAG}_lines.txt; cat gpurun_out/${TAG}_stalls.txt
