set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_host_cli.py -m gpu -q ) > gpurun_out/r2k_tests.log 2>&1; tail -4 gpurun_out/r2k_tests.log
timeout 900 python bench.py > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err; tail -c 400 gpurun_out/r2k_bench_n1.err
timeout 900 python bench.py --impl reference > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err; tail -c 400 gpurun_out/r2k_bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; tail -2 gpurun_out/r2k_smoke.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2k_bench_n1.json"))
print("ms/step", d["ms_per_step"], "Gb/s", d["value"]/1e9, "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e9, "ascii", d["e2e_ascii"]["ms_per_step"])
print("roofline", d["roofline"])
print("parity", d["parity_check"]); print("cpu", d["cpu_baseline"])
for k,v in d["kernels"].items(): print(f"  {k:18s} x{v['launches_per_step']:.0f} {v['ms_per_launch']:.4f} {v['bound']['bound']:5s} frac {v['bound']['frac']:.3f}  hbm_view {v['hbm_view']['frac_of_hbm']:.3f}")
x=d["extra"]["configs[2]"]; print("configs[2]", x["ms_per_step"], x["value"]/1e9, x["e2e"]["ms_per_step"])
r=json.load(open("gpurun_out/r2k_bench_ref.json")); print("reference", r["value"]/1e9, r["ms_per_step"], r["cpu_baseline"]["cores"])
PY
