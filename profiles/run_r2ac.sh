set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2ac_bigset.log
for pct in 50 100; do
  BRGPU_COMPACT_MAX_PCT=$pct timeout 600 python profiles/bigset_probe.py shard 100 50 0.12 8 >> gpurun_out/r2ac_bigset.log 2>> gpurun_out/r2ac_bigset.err
done
BRGPU_COMPACT_MAX_PCT=100 BRGPU_NO_POS8=1 timeout 600 python profiles/bigset_probe.py shard 100 50 0.12 8 >> gpurun_out/r2ac_bigset.log 2>> gpurun_out/r2ac_bigset.err
cat gpurun_out/r2ac_bigset.log; tail -5 gpurun_out/r2ac_bigset.err
