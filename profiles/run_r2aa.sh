set -x
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2aa_bench_n2.json 2> gpurun_out/r2aa_bench_n2.err; tail -c 800 gpurun_out/r2aa_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2aa_bench_n2.json") if l.startswith("{")][-1])
print("N=2 ms/step", d["ms_per_step"], d["value"]/1e9, "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e9, {k:v for k,v in d["parity_check"].items() if k!="bitfield_blake2b"})
x=d["extra"]["configs[3]"]; print("configs[3]", x["ms_per_step"], x["value"]/1e9, (x.get("e2e") or {}).get("ms_per_step"), {k:v for k,v in (x.get("parity_check") or {}).items() if k!="bitfield_blake2b"})
for k,v in x["kernels"].items(): print(f"  {k:20s} x{v['launches_per_step']:.0f} {v['ms_per_launch']:.4f}")
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 profiles/multi_gpu_phases.py weak > gpurun_out/r2aa_phases_n2.log 2>&1; tail -14 gpurun_out/r2aa_phases_n2.log
