#!/bin/bash
# end of round 2: launch list of the bench command, --set full captures of the env kernels after the last changes
# (headline one-warp kernel, role-split kernel at 8,192 worlds, fused self-play rollout, Balance-Beam)
set -u
mkdir -p gpurun_out /tmp/cub && (cd /tmp/cub && cuobjdump -xelf all $OLDPWD/diverse_conventions_b200/libocb.so > /dev/null 2>&1)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2g_launches_bench.csv \
  python bench.py --steps 10 --warmup 3 --e2e-steps 10 --no-cpu-baseline --policy-T 20 > gpurun_out/r2g_ncu_bench.log 2>&1; tail -c 200 gpurun_out/r2g_ncu_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oc_rollout -s 3 -c 1 -o gpurun_out/r2g_oc_rollout_full -f \
  python bench.py --steps 6 --warmup 3 --e2e-steps 10 --no-cpu-baseline --no-config4 --no-config5 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oc_rollout_split -s 3 -c 1 -o gpurun_out/r2g_oc_rollout_split_full -f \
  python tools/sweep.py --layouts simple --worlds 8192 --lanes 16 --quick --tma 1 --passes 3 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_fused -s 1 -c 1 -o gpurun_out/r2g_rollout_fused_full -f \
  python tools/rollout_bench.py --mode selfplay --layouts simple --worlds 8192 --T 100 --graph 0 --iters 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bb_kernel -s 8 -c 1 -o gpurun_out/r2g_bb_full -f \
  python tools/bb_bench.py --worlds 65536 --iters 5 > /dev/null 2>&1
for r in oc_rollout oc_rollout_split rollout_fused bb; do
  python tools/ncu_summary.py gpurun_out/r2g_${r}_full.ncu-rep gpurun_out/r2g_ncu_full_${r}_summary.json > /dev/null
done
python tools/ncu_lines_by_source.py gpurun_out/r2g_oc_rollout_full.ncu-rep /tmp/cub/oc_kernels.sm_100a.cubin oc_rollout_kernelILi2ELi1ELb0E 40 > gpurun_out/r2g_lines_oc_rollout.txt 2>&1
python tools/ncu_lines_by_source.py gpurun_out/r2g_oc_rollout_split_full.ncu-rep /tmp/cub/oc_kernels.sm_100a.cubin oc_rollout_split_kernelILi2ELi1E 40 > gpurun_out/r2g_lines_oc_rollout_split.txt 2>&1
python tools/ncu_lines_by_source.py gpurun_out/r2g_bb_full.ncu-rep /tmp/cub/bb_kernels.sm_100a.cubin 9bb_kernelE 40 > gpurun_out/r2g_lines_bb.txt 2>&1
for r in oc_rollout oc_rollout_split bb; do
ncu -i gpurun_out/r2g_${r}_full.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; units=rows[1]; vals=rows[2]
for h,u,v in zip(hdr,units,vals):
    if ('issue_stalled' in h and 'per_warp_active.pct' in h) or h in ('smsp__inst_executed.sum','gpu__time_duration.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes_write.sum','dram__bytes_read.sum','sm__warps_active.avg.pct_of_peak_sustained_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'): print(h,u,v)
" > gpurun_out/r2g_stalls_${r}.txt
done
rm -f gpurun_out/r2g_*_full.ncu-rep
head -12 gpurun_out/r2g_lines_oc_rollout_split.txt; cat gpurun_out/r2g_stalls_oc_rollout_split.txt | head -40; grep -c . gpurun_out/r2g_launches_bench.csv
echo done
