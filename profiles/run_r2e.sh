set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_correct.py -m gpu -q -x ) > gpurun_out/r2e_tests.log 2>&1; tail -5 gpurun_out/r2e_tests.log
timeout 300 python bench.py --no-extra --no-parity --no-cpu-baseline > gpurun_out/r2e_bench_default.json 2> gpurun_out/r2e_bench_default.err
BRGPU_SCAN=groups timeout 300 python bench.py --no-extra --no-parity --no-cpu-baseline > gpurun_out/r2e_bench_groups.json 2> gpurun_out/r2e_bench_groups.err
BRGPU_SCAN=warp timeout 300 python bench.py --no-extra --no-parity --no-cpu-baseline > gpurun_out/r2e_bench_warp.json 2> gpurun_out/r2e_bench_warp.err
python - <<'PY'
import json
for v in ("default","groups","warp"):
    try:
        d=json.load(open(f"gpurun_out/r2e_bench_{v}.json"))
        print(v, round(d["ms_per_step"],3), {k:x["ms_per_launch"] for k,x in d["kernels"].items() if k.startswith("scan_") or k.startswith("merge_")})
    except Exception as e:
        print(v, "ERR", e)
PY
