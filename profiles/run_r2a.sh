set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_tests.log 2>&1; tail -5 gpurun_out/r2a_tests.log
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench.err; head -c 1500 gpurun_out/r2a_bench.json
for v in mb12 mb6; do BRGPU_LIBRARY=$PWD/br_b200/libbrgpu_$v.so timeout 300 python bench.py --no-extra --no-parity --no-cpu-baseline > gpurun_out/r2a_bench_$v.json 2> gpurun_out/r2a_bench_$v.err; done
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,launch__registers_per_thread
timeout 600 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/ncu_r2_counters_one_two.csv --metrics $M python profiles/ncu_step.py one two > gpurun_out/ncu_a.log 2>&1
timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/ncu_r2_counters_graph_greedy_gap.csv --metrics $M python profiles/ncu_step.py graph greedy gap_size > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_a.log gpurun_out/ncu_b.log
