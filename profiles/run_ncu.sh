#!/bin/bash
# Profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then one
# `--set full` capture of the hot kernels.  Run under gpurun from the repo root:
#   gpurun --timeout 1500 -- 'bash profiles/run_ncu.sh r1'
# Outputs land in gpurun_out/ (scratch); summaries are copied into profiles/ by hand.
set -u
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e-pipeline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'coarse_hist_kernel|coarse_scatter_kernel|fine_partition_kernel|bucket_count_kernel|compact_blocks_kernel|solid_bitmap_kernel|scan_spec8_kernel|scan_spec_kernel|scan_merge_kernel|reverse_slots_kernel' -c 23 \
    -o /tmp/prof_${TAG} -f $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
# the report itself is too large to travel back (gpurun merges <= 64 MiB): summarise it here
NCU_REP_DIR=/tmp SUMMARY_OUT=gpurun_out python profiles/summarize.py ${TAG} > /dev/null
for kern in scan_spec8_kernel scan_spec_kernel solid_bitmap_kernel bucket_count_kernel coarse_scatter_kernel fine_partition_kernel; do
    python profiles/source_hotspots.py /tmp/prof_${TAG}.ncu-rep $kern 0 40 > gpurun_out/hotspots_${TAG}_${kern}.txt 2>/dev/null
done
tail -3 gpurun_out/plain_${TAG}.log
