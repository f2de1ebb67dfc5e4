set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2al_bench_n1.json 2> gpurun_out/r2al_bench_n1.err; tail -c 300 gpurun_out/r2al_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2al_bench_n1.json") if l.startswith("{")][-1])
print("N=1 ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], {k:v for k,v in d["parity_check"].items() if k!="bitfield_blake2b"})
print("   ", " ".join(f"{k}={v['ms_per_launch']:.3f}" for k,v in d["kernels"].items() if k.startswith(("scan","merge","solid"))))
x=d["extra"]["configs[2]"]; print("configs[2]", x["ms_per_step"], x["value"]/1e9, x["e2e"]["ms_per_step"], {k:v for k,v in x["parity_check"].items() if k not in ("bitfield_blake2b","note")})
print("   ", " ".join(f"{k}={v['ms_per_launch']:.3f}" for k,v in x["kernels"].items() if k.startswith(("scan","merge","solid"))))
PY
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/r2al_smoke.log 2>&1; tail -2 gpurun_out/r2al_smoke.log
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2al_tests.log 2>&1; tail -6 gpurun_out/r2al_tests.log
