#!/bin/bash
# Second profiling recipe: only the `--set full` capture, for a chosen kernel regex.
#   gpurun --timeout 1500 -- 'bash profiles/run_ncu2.sh r1g "bucket_count_kernel|bucket_scatter_kernel|bucket_hist_kernel" 6'
set -u
TAG=${1:-r1}
REGEX=${2:-scan_kernel}
COUNT=${3:-6}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -c $COUNT \
    -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
tail -c 600 gpurun_out/plain_${TAG}.log
