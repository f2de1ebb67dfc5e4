set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s ) > gpurun_out/r2t_multi.log 2>&1; tail -16 gpurun_out/r2t_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-extra > gpurun_out/r2t_bench_n2.json 2> gpurun_out/r2t_bench_n2.err; tail -c 1500 gpurun_out/r2t_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2t_bench_n2.json") if l.startswith("{")][-1])
print("N=2 ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d.get("parity_check"))
tot=sum(v['ms_per_launch']*v['launches_per_step'] for v in d['kernels'].values()); print("kernel sum", tot)
for k,v in d["kernels"].items(): print(f"  {k:20s} x{v['launches_per_step']:.0f} {v['ms_per_launch']:.4f}")
PY
