set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2ah_tests.log 2>&1; tail -6 gpurun_out/r2ah_tests.log
timeout 900 python bench.py > gpurun_out/r2ah_bench_n1.json 2> gpurun_out/r2ah_bench_n1.err; tail -c 300 gpurun_out/r2ah_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2ah_bench_n1.json") if l.startswith("{")][-1])
print("N=1 ms/step", d["ms_per_step"], "value", d["value"]/1e9, "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e9, "ascii", d["e2e_ascii"]["ms_per_step"])
print({k:v for k,v in d["parity_check"].items() if k!="bitfield_blake2b"})
x=d["extra"]["configs[2]"]; print("configs[2]", x["ms_per_step"], x["value"]/1e9, x["e2e"]["ms_per_step"], x.get("parity_check"))
PY
