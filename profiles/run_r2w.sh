set -x
mkdir -p gpurun_out
for v in base el cg; do
  lib=br_b200/libbrgpu.so; [ $v != base ] && lib=br_b200/libbrgpu_$v.so
  BRGPU_LIBRARY=$PWD/$lib timeout 600 python bench.py --steps 5 --warmup 3 --no-extra > gpurun_out/r2w_bench_$v.json 2> gpurun_out/r2w_bench_$v.err; tail -c 300 gpurun_out/r2w_bench_$v.err
done
python - <<'PY'
import json
for tag in ("base","el","cg"):
    d=json.loads([l for l in open(f"gpurun_out/r2w_bench_{tag}.json") if l.startswith("{")][-1])
    print(tag, "ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], {k:v for k,v in (d.get("parity_check") or {}).items() if k!='bitfield_blake2b'})
    print("   ", " ".join(f"{k}={v['ms_per_launch']:.3f}" for k,v in d["kernels"].items() if k.startswith(("solid","scan","merge"))))
PY
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/r2w_group1gpu.log 2>&1; tail -3 gpurun_out/r2w_group1gpu.log
