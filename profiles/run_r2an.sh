set -x
mkdir -p gpurun_out
for v in default groups; do
  if [ $v = default ]; then unset BRGPU_SCAN; else export BRGPU_SCAN=$v; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-extra --no-parity > gpurun_out/r2an_bench_$v.json 2> gpurun_out/r2an_bench_$v.err
done
python - <<'PY'
import json
for tag in ("default","groups"):
    d=json.loads([l for l in open(f"gpurun_out/r2an_bench_{tag}.json") if l.startswith("{")][-1])
    print(tag, "ms/step", d["ms_per_step"], " ".join(f"{k}={v['ms_per_launch']:.3f}" for k,v in d["kernels"].items() if k.startswith(("scan","merge"))))
PY
