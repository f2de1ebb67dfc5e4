set -x
mkdir -p gpurun_out
# A/B: compacted set pinned in L2 (default) vs not
python - <<'PY'
import torch
p=torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size)
import ctypes
rt=ctypes.CDLL("libcudart.so.12")
v=ctypes.c_int()
for name,a in (("persistingL2CacheMaxSize",108),("accessPolicyMaxWindowSize",109)):
    rt.cudaDeviceGetAttribute(ctypes.byref(v),a,0); print(name,v.value)
PY
( timeout 900 python -m pytest tests/test_gpu_correct.py tests/test_gpu_set.py -m gpu -q -x ) > gpurun_out/r2p_tests.log 2>&1; tail -5 gpurun_out/r2p_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-extra --no-parity > gpurun_out/r2p_bench_pin.json 2> gpurun_out/r2p_bench_pin.err; tail -c 600 gpurun_out/r2p_bench_pin.err
BRGPU_NO_L2_PERSIST=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-extra --no-parity > gpurun_out/r2p_bench_nopin.json 2> gpurun_out/r2p_bench_nopin.err; tail -c 600 gpurun_out/r2p_bench_nopin.err
python - <<'PY'
import json
for tag in ("pin","nopin"):
    d=json.loads([l for l in open(f"gpurun_out/r2p_bench_{tag}.json") if l.startswith("{")][-1])
    print(tag, "ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
    for k,v in d["kernels"].items(): print(f"  {k:20s} x{v['launches_per_step']:.0f} {v['ms_per_launch']:.4f}")
PY
