set -u
mkdir -p /tmp/cub && (cd /tmp/cub && cuobjdump -xelf all $OLDPWD/diverse_conventions_b200/libocb.so > /dev/null 2>&1)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oc_rollout -s 3 -c 1 -o gpurun_out/r2h_oc_rollout_full -f python bench.py --steps 6 --warmup 3 --e2e-steps 10 --no-cpu-baseline --no-config4 --no-config5 > /dev/null 2>&1
python tools/ncu_lines_by_source.py gpurun_out/r2h_oc_rollout_full.ncu-rep /tmp/cub/oc_kernels.sm_100a.cubin oc_rollout_kernelILi2ELi1ELb0E 45 > gpurun_out/r2h_lines_oc_rollout.txt 2>&1
ncu -i gpurun_out/r2h_oc_rollout_full.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; units=rows[1]; vals=rows[2]
for h,u,v in zip(hdr,units,vals):
    if ('issue_stalled' in h and 'per_warp_active.pct' in h) or h in ('smsp__inst_executed.sum','gpu__time_duration.sum','smsp__issue_active.avg.pct_of_peak_sustained_active'): print(h,u,v)
" > gpurun_out/r2h_stalls_oc_rollout.txt
rm -f gpurun_out/r2h_oc_rollout_full.ncu-rep
cat gpurun_out/r2h_lines_oc_rollout.txt | head -50; cat gpurun_out/r2h_stalls_oc_rollout.txt
