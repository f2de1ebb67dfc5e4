set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2ak_bench_n1.json 2> gpurun_out/r2ak_bench_n1.err; tail -c 300 gpurun_out/r2ak_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2ak_bench_n1.json") if l.startswith("{")][-1])
print("N=1 ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], {k:v for k,v in d["parity_check"].items() if k!="bitfield_blake2b"})
x=d["extra"]["configs[2]"]; print("configs[2]", x["ms_per_step"], x["value"]/1e9, x["e2e"]["ms_per_step"], {k:v for k,v in x["parity_check"].items() if k not in ("bitfield_blake2b","note")})
print("   ", " ".join(f"{k}={v['ms_per_launch']:.3f}" for k,v in x["kernels"].items() if k.startswith(("scan","merge"))))
PY
