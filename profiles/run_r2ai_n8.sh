set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2ai_bench_n8.json 2> gpurun_out/r2ai_bench_n8.err; tail -c 1500 gpurun_out/r2ai_bench_n8.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2ai_bench_n8.json") if l.startswith("{")][-1])
print("N=8 ms/step", d["ms_per_step"], "value", d["value"]/1e9, "e2e", d["e2e"]["ms_per_step"], d.get("e2e_ascii",{}).get("ms_per_step"))
print(d.get("parity_check"))
for k,v in d["kernels"].items(): print(f"  {k:20s} x{v['launches_per_step']:.0f} {v['ms_per_launch']:.4f}")
for leg in ("configs[3]","configs[4]"):
    x=d.get("extra",{}).get(leg)
    if x:
        print(leg, "ms/step", x["ms_per_step"], "Gbases/s", x["value"]/1e9, (x.get("e2e") or {}).get("ms_per_step"), x.get("parity_check"))
        for k,v in x["kernels"].items(): print(f"  {k:20s} x{v['launches_per_step']:.0f} {v['ms_per_launch']:.4f}")
PY
