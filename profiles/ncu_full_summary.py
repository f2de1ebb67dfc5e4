#!/usr/bin/env python3
"""`ncu -i X.ncu-rep --page raw --csv` -> the per-launch metric digest kept as profiles/ncu_*final.txt.

    python profiles/ncu_full_summary.py gpurun_out/ncu_r2final_raw.csv > profiles/ncu_r2final.txt
"""
import csv
import sys

METRICS = """gpu__time_duration.sum smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active
sm__warps_active.avg.pct_of_peak_sustained_active launch__registers_per_thread launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem dram__bytes_read.sum dram__bytes_write.sum lts__t_sector_hit_rate.pct
lts__throughput.avg.pct_of_peak_sustained_elapsed l1tex__t_sector_hit_rate.pct
smsp__thread_inst_executed_per_inst_executed.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio""".split()


def main(path):
    rows = list(csv.reader(open(path, newline="")))
    head, units = rows[0], rows[1]
    col = {}
    for i, h in enumerate(head):  # a metric may appear under several section prefixes: first non-empty wins per row
        col.setdefault(h.split(".", 2)[-1] if h.count(".") > 3 and h.split(".")[0].isupper() else h, []).append(i)
    print("# ncu --set full --clock-control none --import-source on, one step of BASELINE configs[1] (profiles/run_ncu_r2_full.sh r2final ...)")
    print("# launches in order of the step; times under ncu are cold-cache and serialised (shares, not absolutes, carry over)")
    name_i = head.index("Kernel Name")
    for n, r in enumerate(rows[2:]):
        print(f"\n== launch {n}: {r[name_i][:110]}")
        for m in METRICS:
            idx = [i for h, ii in col.items() if h == m or h.endswith("." + m) for i in ii]
            val = next(((r[i], units[i]) for i in idx if i < len(r) and r[i] != ""), None)
            if val:
                print(f"  {m:86s} {val[0]:>18s} {val[1]}")


if __name__ == "__main__":
    main(sys.argv[1])
