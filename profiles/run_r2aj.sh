set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2aj_bigset.log
timeout 600 python profiles/bigset_probe.py shard 1000 30 0.10 24 1 >> gpurun_out/r2aj_bigset.log 2>> gpurun_out/r2aj_bigset.err
BRGPU_KEEP_SUMMARY=1 timeout 600 python profiles/bigset_probe.py shard 1000 30 0.10 24 1 >> gpurun_out/r2aj_bigset.log 2>> gpurun_out/r2aj_bigset.err
cat gpurun_out/r2aj_bigset.log; tail -5 gpurun_out/r2aj_bigset.err
( timeout 600 python -m pytest tests/test_gpu_correct.py -m gpu -q -x -k "large_k" ) > gpurun_out/r2aj_tests.log 2>&1; tail -3 gpurun_out/r2aj_tests.log
