set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_correct.py tests/test_gpu_set.py tests/test_gpu_scale.py -m gpu -q -x ) > gpurun_out/r2s_tests.log 2>&1; tail -5 gpurun_out/r2s_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-extra > gpurun_out/r2s_bench_pos8.json 2> gpurun_out/r2s_bench_pos8.err; tail -c 600 gpurun_out/r2s_bench_pos8.err
BRGPU_NO_POS8=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-extra --no-parity > gpurun_out/r2s_bench_nopos8.json 2> gpurun_out/r2s_bench_nopos8.err; tail -c 600 gpurun_out/r2s_bench_nopos8.err
python - <<'PY'
import json
for tag in ("pos8","nopos8"):
    d=json.loads([l for l in open(f"gpurun_out/r2s_bench_{tag}.json") if l.startswith("{")][-1])
    print(tag, "ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d.get("parity_check"))
    for k,v in d["kernels"].items(): print(f"  {k:20s} x{v['launches_per_step']:.0f} {v['ms_per_launch']:.4f}")
PY
