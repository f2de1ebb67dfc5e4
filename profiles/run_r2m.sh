set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_correct.py -m gpu -q -x ) > gpurun_out/r2m_tests.log 2>&1; tail -3 gpurun_out/r2m_tests.log
for r in 1 2; do timeout 300 python bench.py --no-extra --no-parity --no-cpu-baseline > gpurun_out/r2m_bench_$r.json 2> gpurun_out/r2m_bench_$r.err; done
python - <<'PY'
import json
for v in "12":
    try:
        d=json.load(open(f"gpurun_out/r2m_bench_{v}.json"))
        print(v, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), {k:x["ms_per_launch"] for k,x in d["kernels"].items() if k.startswith("scan_") or k.startswith("merge_")})
    except Exception as e:
        print(v, "ERR", e, open(f"gpurun_out/r2m_bench_{v}.err").read()[-1500:])
PY
