#!/usr/bin/env python3
"""Turn the scratch ncu outputs of profiles/run_ncu.sh into small committed summaries.

    python profiles/summarize.py r1a      # reads gpurun_out/launches_r1a.csv, gpurun_out/prof_r1a.ncu-rep
writes profiles/launches_<tag>.txt (per-kernel share of the step) and profiles/ncu_<tag>.txt
(the counters DESIGN.md argues from, per profiled launch).
"""
import collections
import csv
import subprocess
import sys
from pathlib import Path

import os

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1]
# SUMMARY_OUT: where the summaries go (default profiles/; on the GPU box gpurun_out/, so that only
# the small text files travel back); NCU_REP_DIR: where prof_<tag>.ncu-rep lives
out_dir = Path(os.environ.get("SUMMARY_OUT", ROOT / "profiles"))

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__sectors_read.sum",
    "dram__sectors_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_requests_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_requests_srcunit_tex_op_atom_dot_cas.sum", "lts__t_sectors_srcunit_tex_op_atom.sum",
    "lts__t_sectors_srcunit_tex_op_red.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers",
]

launch_csv = ROOT / "gpurun_out" / f"launches_{tag}.csv"
if launch_csv.exists():
    rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(out_dir / f"launches_{tag}.txt", "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none ({launch_csv.name}); cold-cache, serialised:\n")
        f.write("# compare SHARES with bench.py's live CUDA-event shares, not absolutes\n")
        f.write(f"{'kernel':34s} {'launches':>8s} {'total_ms':>10s} {'ms/launch':>10s} {'share':>7s}\n")
        for n, (c, v) in agg.items():
            f.write(f"{n:34s} {c:8d} {v / 1e6:10.3f} {v / 1e6 / c:10.4f} {v / tot:7.1%}\n")
    print((out_dir / f"launches_{tag}.txt").read_text())

rep = Path(os.environ.get("NCU_REP_DIR", ROOT / "gpurun_out")) / f"prof_{tag}.ncu-rep"
if rep.exists():
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out_dir / f"ncu_{tag}.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ({rep.name}), one block per profiled launch\n")
        for d in data:
            f.write(f"\n== {d[idx['Kernel Name']].split('(')[0]}  (launch id {d[idx['ID']]})\n")
            for m in METRICS:
                if m in idx and d[idx[m]] not in ("", "n/a"):
                    f.write(f"  {m:78s} {d[idx[m]]:>18s} {units[idx[m]]}\n")
    print("wrote", out_dir / f"ncu_{tag}.txt")

    # DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) under bench.py's kernel
    # names -> profiles/traffic.json, which bench.py reports as roofline.traffic
    import json
    import re

    def bench_name(kernel):
        """ncu's kernel name -> the name bench.py reports (scan_spec8_kernel<0, 17> -> scan_one ...)."""
        m = re.match(r"(?:void )?(?:brgpu::)?(\w+?)(?:_kernel)?(?:<([^>]*)>)?$", kernel.strip())
        if not m:
            return kernel
        base = m.group(1)
        args = [a.strip().replace("(int)", "") for a in (m.group(2) or "").split(",") if a.strip()]
        methods = ["one", "two", "graph", "greedy", "gap_size"]
        if base in ("scan_spec", "scan_spec8") and args:
            return "scan_" + methods[int(args[0])]
        if base == "scan_merge" and args:
            return "merge_" + methods[int(args[0])]
        return base

    def to_bytes(val, unit):
        v = float(val.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)

    acc = collections.OrderedDict()
    for d in data:
        name = bench_name(d[idx["Kernel Name"]].split("(")[0])
        rd = to_bytes(d[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(d[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        ms = float(d[idx["gpu__time_duration.sum"]].replace(",", "")) * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(
            units[idx["gpu__time_duration.sum"]], 1.0)
        a = acc.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "ms": 0.0})
        a["launches"] += 1
        a["dram_bytes"] += rd + wr
        a["ms"] += ms
    traffic = {"source": f"ncu --set full --clock-control none ({rep.name}); mean over the captured launches of one step",
               "kernels": {n: {"dram_bytes_per_launch": round(a["dram_bytes"] / a["launches"]),
                               "launches_captured": a["launches"], "ms_per_launch_under_ncu": round(a["ms"] / a["launches"], 4)}
                           for n, a in acc.items()}}
    (out_dir / "traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
    print("wrote", out_dir / "traffic.json")
