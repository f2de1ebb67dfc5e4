set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2y_tests.log 2>&1; tail -6 gpurun_out/r2y_tests.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,launch__registers_per_thread
timeout 600 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/ncu_r2_counters_one_two.csv --metrics $M python profiles/ncu_step.py one two > gpurun_out/ncu_a.log 2>&1
timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/ncu_r2_counters_graph_greedy_gap.csv --metrics $M python profiles/ncu_step.py graph greedy gap_size > gpurun_out/ncu_b.log 2>&1
bash profiles/run_ncu_r2_full.sh r2final "scan_spec8_kernel|scan_spec_kernel|solid_bitmap_kernel|bucket_count_kernel|fine_partition_kernel|coarse_scatter_kernel|coarse_hist_kernel|scan_merge_kernel" 12 one two
ncu -i gpurun_out/prof_r2final.ncu-rep --page raw --csv > gpurun_out/ncu_r2final_raw.csv 2>/dev/null
for kk in scan_spec8_kernel scan_spec_kernel solid_bitmap_kernel bucket_count_kernel fine_partition_kernel coarse_scatter_kernel; do python profiles/source_hotspots.py gpurun_out/prof_r2final.ncu-rep $kk 0 40 > gpurun_out/hotspots_r2final_$kk.txt 2>&1; done
