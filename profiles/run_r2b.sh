set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_hash.py tests/test_gpu_synth.py tests/test_host_cli.py "tests/test_gpu_set.py::test_set_built_from_a_stream_of_chunks_equals_the_one_shot_set" -m gpu -x -q ) > gpurun_out/r2b_tests.log 2>&1; tail -25 gpurun_out/r2b_tests.log
for v in s8x10 s8x12; do BRGPU_LIBRARY=$PWD/br_b200/libbrgpu_$v.so timeout 300 python bench.py --no-extra --no-parity --no-cpu-baseline > gpurun_out/r2b_bench_$v.json 2> gpurun_out/r2b_bench_$v.err; done
timeout 300 python bench.py --no-extra --no-parity --no-cpu-baseline > gpurun_out/r2b_bench_base.json 2> gpurun_out/r2b_bench_base.err
python - <<'PY'
import json
for v in ("base","s8x10","s8x12"):
    try:
        d=json.load(open(f"gpurun_out/r2b_bench_{v}.json"))
        print(v, round(d["ms_per_step"],3), {k:x["ms_per_launch"] for k,x in d["kernels"].items() if k.startswith("scan_")}, d["roofline"])
    except Exception as e:
        print(v, "ERR", e)
PY
