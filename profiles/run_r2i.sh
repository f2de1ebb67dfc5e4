set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_set.py tests/test_gpu_scale.py -m gpu -q -x ) > gpurun_out/r2i_tests.log 2>&1; tail -5 gpurun_out/r2i_tests.log
for v in 0 1; do BRGPU_COUNT_BLOCK_ONLY=$v timeout 300 python bench.py --no-extra --no-parity --no-cpu-baseline > gpurun_out/r2i_bench_$v.json 2> gpurun_out/r2i_bench_$v.err; done
python - <<'PY'
import json
for v in "01":
    try:
        d=json.load(open(f"gpurun_out/r2i_bench_{v}.json"))
        print(v, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), {k:x["ms_per_launch"] for k,x in d["kernels"].items() if k in ("coarse_hist","coarse_scatter","fine_partition","bucket_count","compact_blocks","summary_popc")})
    except Exception as e:
        print(v, "ERR", e, open(f"gpurun_out/r2i_bench_{v}.err").read()[-1500:])
PY
