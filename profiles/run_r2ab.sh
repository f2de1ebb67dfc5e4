set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2ab_bigset.log
for v in base anyarm base anyarm; do
  lib=br_b200/libbrgpu.so; [ $v != base ] && lib=br_b200/libbrgpu_$v.so
  BRGPU_LIBRARY=$PWD/$lib timeout 600 python profiles/bigset_probe.py shard 100 50 0.12 8 >> gpurun_out/r2ab_bigset.log 2>> gpurun_out/r2ab_bigset.err
done
cat gpurun_out/r2ab_bigset.log; tail -5 gpurun_out/r2ab_bigset.err
( timeout 900 python -m pytest tests/test_gpu_correct.py tests/test_gpu_scale.py -m gpu -q -x ) > gpurun_out/r2ab_tests.log 2>&1; tail -3 gpurun_out/r2ab_tests.log
