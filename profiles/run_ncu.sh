#!/bin/bash
# Profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then one
# `--set full` capture of the hot kernels.  Run under gpurun from the repo root:
#   gpurun --timeout 1500 -- 'bash profiles/run_ncu.sh r1'
# Outputs land in gpurun_out/ (scratch); summaries are copied into profiles/ by hand.
set -u
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'count_kernel|solid_bitmap_kernel|scan_kernel|spectrum_threshold_kernel' -c 10 \
    -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/plain_${TAG}.log
