set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; tail -c 400 gpurun_out/r2z_bench_n1.err
timeout 900 python bench.py --impl reference > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; tail -c 400 gpurun_out/r2z_bench_ref.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2z_bench_n1.json") if l.startswith("{")][-1])
print("N=1 ms/step", d["ms_per_step"], "value", d["value"]/1e9, "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e9, "ascii", d["e2e_ascii"]["ms_per_step"])
print({k:v for k,v in d["parity_check"].items() if k!="bitfield_blake2b"})
print("roofline", d["roofline"]); print("aggregates", d["aggregates"]); print("cpu_baseline", d["cpu_baseline"]); print("clocks", d["clocks"], "launches", d["gpu_launches"])
for k,v in d["kernels"].items(): print(f"  {k:20s} x{v['launches_per_step']:.0f} {v['ms_per_launch']:.4f} {v['bound']['bound']} {v['bound']['frac']}")
x=d["extra"]["configs[2]"]; print("configs[2]", x["ms_per_step"], x["value"]/1e9, x["e2e"]["ms_per_step"], {k:v for k,v in (x.get("parity_check") or {}).items() if k!="bitfield_blake2b"})
r=json.loads([l for l in open("gpurun_out/r2z_bench_ref.json") if l.startswith("{")][-1]); print("reference", r["value"]/1e9, r["ms_per_step"], r["cpu_baseline"])
PY
