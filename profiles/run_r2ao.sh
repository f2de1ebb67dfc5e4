set -x
mkdir -p gpurun_out
( timeout 150 python -m pytest tests/test_host_cli.py tests/test_gpu_multi.py -m gpu -q -x ) > gpurun_out/r2ao_tests.log 2>&1; tail -5 gpurun_out/r2ao_tests.log
