set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2ag_bigset.log
for i in 1; do
timeout 600 python profiles/bigset_probe.py shard 100 50 0.12 8 >> gpurun_out/r2ag_bigset.log 2>> gpurun_out/r2ag_bigset.err
BRGPU_FINE_IN_SCANS=1 timeout 600 python profiles/bigset_probe.py shard 100 50 0.12 8 >> gpurun_out/r2ag_bigset.log 2>> gpurun_out/r2ag_bigset.err
BRGPU_NO_FINE_SUMMARY=1 timeout 600 python profiles/bigset_probe.py shard 100 50 0.12 8 >> gpurun_out/r2ag_bigset.log 2>> gpurun_out/r2ag_bigset.err
done
timeout 600 python profiles/bigset_probe.py shard 1000 30 0.10 24 1 >> gpurun_out/r2ag_bigset.log 2>> gpurun_out/r2ag_bigset.err
cat gpurun_out/r2ag_bigset.log; tail -5 gpurun_out/r2ag_bigset.err
