#!/bin/bash
# `ncu --set full` capture of chosen kernels of one configs[1] step (profiles/ncu_step.py):
#   gpurun --timeout 1500 -- 'bash profiles/run_ncu_r2_full.sh r2scan "scan_spec8_kernel|scan_spec_kernel" 2 one two'
set -u
TAG=${1:-r2}
REGEX=${2:-scan_spec}
COUNT=${3:-2}
shift 3
mkdir -p gpurun_out
python profiles/ncu_step.py "$@" > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$REGEX" -c $COUNT \
    -o gpurun_out/prof_${TAG} -f python profiles/ncu_step.py "$@" > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
tail -c 300 gpurun_out/plain_${TAG}.log
