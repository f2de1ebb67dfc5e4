set -x
mkdir -p gpurun_out
for i in 1 2 3; do ( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s ) > gpurun_out/r2u_multi_$i.log 2>&1; tail -3 gpurun_out/r2u_multi_$i.log; done
