set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s ) > gpurun_out/r2f_multi_parity.log 2>&1; tail -14 gpurun_out/r2f_multi_parity.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --chunk-template-bases 1000000000 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; tail -c 2500 gpurun_out/r2f_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2f_bench_n2.json") if l.startswith("{")][-1])
print("N=2 ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d.get("e2e_ascii",{}).get("ms_per_step"))
print(d.get("parity_check"))
x=d.get("extra",{}).get("configs[3]")
if x:
    print("configs[3] ms/step", x["ms_per_step"], "value", x["value"], x.get("e2e"), x.get("parity_check"), x["config"])
    for k,v in x["kernels"].items(): print(f"  {k:20s} x{v['launches_per_step']:.0f} {v['ms_per_launch']:.4f}")
PY
