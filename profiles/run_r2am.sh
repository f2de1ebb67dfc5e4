set -x
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,launch__registers_per_thread
timeout 300 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/ncu_r2_counters_one_two.csv --metrics $M python profiles/ncu_step.py one two > gpurun_out/ncu_a.log 2>&1
timeout 400 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/ncu_r2_counters_graph_greedy_gap.csv --metrics $M python profiles/ncu_step.py graph greedy gap_size > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_a.log gpurun_out/ncu_b.log
