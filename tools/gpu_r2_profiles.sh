#!/bin/bash
# round 2 ncu evidence (final state): launch list of the bench command, --set full captures of the dominant kernels
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches_bench.csv \
  python bench.py --steps 10 --warmup 3 --e2e-steps 10 --no-cpu-baseline --policy-T 20 > gpurun_out/r2f_ncu_bench.log 2>&1; tail -c 300 gpurun_out/r2f_ncu_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oc_rollout -s 3 -c 2 -o gpurun_out/r2f_oc_rollout_full -f \
  python bench.py --steps 6 --warmup 3 --e2e-steps 10 --no-cpu-baseline --no-config4 --no-config5 > gpurun_out/r2f_ncu_full_oc.log 2>&1; tail -c 200 gpurun_out/r2f_ncu_full_oc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oc_rollout -s 6 -c 1 -o gpurun_out/r2f_step_single_full -f \
  python tools/step_single.py --iters 3 > gpurun_out/r2f_ncu_step_single.log 2>&1; tail -c 200 gpurun_out/r2f_ncu_step_single.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_fused -s 1 -c 1 -o gpurun_out/r2f_rollout_fused_full -f \
  python tools/rollout_bench.py --mode selfplay --layouts simple --worlds 8192 --T 100 --graph 0 --iters 1 > gpurun_out/r2f_ncu_full_fused.log 2>&1; tail -c 300 gpurun_out/r2f_ncu_full_fused.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_fused -s 1 -c 1 -o gpurun_out/r2f_rollout_fused_cross_full -f \
  python tools/rollout_bench.py --mode crossplay --policies 16 --worlds-per-pair 1024 --iters 1 > gpurun_out/r2f_ncu_full_fused_cross.log 2>&1; tail -c 300 gpurun_out/r2f_ncu_full_fused_cross.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:policy_pair -s 2 -c 1 -o gpurun_out/r2f_policy_pair_full -f \
  python tools/policy_bench.py --mode fused --layouts simple --rows 262144 --iters 2 > gpurun_out/r2f_ncu_full_pair.log 2>&1; tail -c 300 gpurun_out/r2f_ncu_full_pair.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:policy_pair -s 2 -c 1 -o gpurun_out/r2f_policy_single_full -f \
  python tools/policy_bench.py --mode act --layouts random1 --rows 524288 --iters 2 > gpurun_out/r2f_ncu_full_single.log 2>&1; tail -c 300 gpurun_out/r2f_ncu_full_single.log
# the reports carry the embedded sources (13-17 MB each) and gpurun brings back at most 64 MiB: summarise here
mkdir -p /tmp/cub && (cd /tmp/cub && cuobjdump -xelf all $OLDPWD/diverse_conventions_b200/libocb.so > /dev/null 2>&1)
for r in oc_rollout step_single rollout_fused rollout_fused_cross policy_pair policy_single; do
  python tools/ncu_summary.py gpurun_out/r2f_${r}_full.ncu-rep gpurun_out/r2f_ncu_full_${r}_summary.json
done
python tools/ncu_lines_by_source.py gpurun_out/r2f_step_single_full.ncu-rep /tmp/cub/oc_kernels.sm_100a.cubin oc_rollout_kernelILi2ELi2ELb1E 25 > gpurun_out/r2f_lines_step_single.txt 2>&1
python tools/ncu_lines_by_source.py gpurun_out/r2f_rollout_fused_cross_full.ncu-rep /tmp/cub/policy_kernels.sm_100a.cubin rollout_fused_kernelILb0ELb0ELi2E 40 > gpurun_out/r2f_lines_fused_cross.txt 2>&1
python tools/ncu_lines_by_source.py gpurun_out/r2f_policy_single_full.ncu-rep /tmp/cub/policy_kernels.sm_100a.cubin policy_pair_kernelILb0E 40 > gpurun_out/r2f_lines_policy_single.txt 2>&1
rm -f gpurun_out/r2f_*_full.ncu-rep
ls -la gpurun_out/
echo done
