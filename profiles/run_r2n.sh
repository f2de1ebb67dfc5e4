set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_synth.py tests/test_gpu_correct.py tests/test_host_cli.py -m gpu -q -x ) > gpurun_out/r2n_tests.log 2>&1; tail -3 gpurun_out/r2n_tests.log
python profiles/host_prep_probe.py 2>&1 | tail -8
timeout 300 python bench.py --no-extra --no-parity --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2n_bench.json"))
print(round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), d["e2e"]["host_clock_ms_per_step"], "ascii", round(d["e2e_ascii"]["ms_per_step"],3), d["e2e_ascii"]["host_clock_ms_per_step"])
PY
