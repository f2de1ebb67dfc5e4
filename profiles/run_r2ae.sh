set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2ae_bigset.log
timeout 600 python profiles/bigset_probe.py shard 100 50 0.12 8 >> gpurun_out/r2ae_bigset.log 2>> gpurun_out/r2ae_bigset.err
BRGPU_NO_FINE_SUMMARY=1 timeout 600 python profiles/bigset_probe.py shard 100 50 0.12 8 >> gpurun_out/r2ae_bigset.log 2>> gpurun_out/r2ae_bigset.err
timeout 600 python profiles/bigset_probe.py shard 1000 30 0.10 24 1 >> gpurun_out/r2ae_bigset.log 2>> gpurun_out/r2ae_bigset.err
BRGPU_NO_FINE_SUMMARY=1 timeout 600 python profiles/bigset_probe.py shard 1000 30 0.10 24 1 >> gpurun_out/r2ae_bigset.log 2>> gpurun_out/r2ae_bigset.err
cat gpurun_out/r2ae_bigset.log; tail -5 gpurun_out/r2ae_bigset.err
