set -x
mkdir -p gpurun_out
for n in 8 4; do
for v in base el; do
  lib=br_b200/libbrgpu.so; [ $v != base ] && lib=br_b200/libbrgpu_$v.so
  BRGPU_LIBRARY=$PWD/$lib timeout 300 python profiles/bigset_probe.py $n >> gpurun_out/r2x_bigset.log 2>> gpurun_out/r2x_bigset.err
done
BRGPU_NO_POS8=1 timeout 300 python profiles/bigset_probe.py $n >> gpurun_out/r2x_bigset.log 2>> gpurun_out/r2x_bigset.err
done
timeout 300 python profiles/bigset_probe.py 1 >> gpurun_out/r2x_bigset.log 2>> gpurun_out/r2x_bigset.err
cat gpurun_out/r2x_bigset.log; tail -5 gpurun_out/r2x_bigset.err
