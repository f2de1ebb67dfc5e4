"""ncu CSV (profiles/ncu_step.py under `ncu --csv --metrics ...`) -> profiles/kernel_counters.json.

    python profiles/summarize_r2.py gpurun_out/ncu_r2_counters_one_two.csv [more.csv ...] > profiles/launches_r2.txt

Kernel functions are mapped onto the names bench.py's per-kernel table uses (a name = one ProfScope of
the library; the reversed pass carries the suffix _rev: orientation = parity of the reverse_slots
launches seen so far in the step).  Per name: launches, warp instructions, DRAM bytes and time per launch.
The launch list with per-launch times goes to stdout (kept as profiles/launches_r2.txt).
"""
import csv
import json
import re
import sys
from collections import OrderedDict
from pathlib import Path

METHODS = ["one", "two", "graph", "greedy", "gap_size"]
PRIMARY = {  # kernel function -> (bench name, opens a scope)
    "coarse_hist_kernel": "coarse_hist", "coarse_scatter_kernel": "coarse_scatter", "fine_partition_kernel": "fine_partition",
    "bucket_count_kernel": "bucket_count", "summary_popc_kernel": "summary_popc", "compact_blocks_kernel": "compact_blocks",
    "compact_blocks_stream_kernel": "compact_blocks", "block_bytes_kernel": "block_bytes", "peer_pull_kernel": "peer_pull", "reverse_slots_kernel": "reverse_slots", "seg_count_kernel": "seg_count",
    "bucket_hist_kernel": "bucket_hist", "bucket_scatter_kernel": "bucket_scatter", "build_summary_kernel": "build_summary",
    "count_kernel": "count_kmers", "spectrum_threshold_kernel": "spectrum_threshold",
}
SECONDARY = {"bucket_cursor_kernel": "coarse_scatter", "scan_tile_sums_kernel": "exclusive_scan", "scan_tile_bases_kernel": None,
             "scan_apply_kernel": None}


def parse(path):
    """[(kernel function with template args, {metric: value})] in launch order."""
    rows = OrderedDict()
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        i = int(r["ID"])
        rows.setdefault(i, (r["Kernel Name"], {}))
        try:
            rows[i][1][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            pass
    return list(rows.values())


def bench_name(fn, state):
    base = re.sub(r"^(void )?(brgpu::)?((fast|cnt)::)?", "", fn.split("(")[0].strip())
    m = re.match(r"(\w+?)(<(.*)>)?$", base)
    if not m:
        return None, False
    name, targs = m.group(1), (m.group(3) or "")
    rev = "_rev" if state["reversals"] % 2 else ""
    if name == "reverse_slots_kernel":
        state["reversals"] += 1
        return "reverse_slots", True
    if name == "solid_bitmap_kernel":
        return "solid_bitmap" + rev, True
    if name in ("scan_spec_kernel", "scan_spec8_kernel"):
        return "scan_" + METHODS[int(targs.split(",")[0])] + rev, True
    if name == "scan_merge_kernel":
        state["merge"] = "merge_" + METHODS[int(targs.split(",")[0])] + rev
        return state["merge"], True
    if name == "scan_splice_kernel":
        return state.get("merge"), False
    if name in PRIMARY:
        return PRIMARY[name], True
    if name in SECONDARY:
        tgt = SECONDARY[name]
        if name == "scan_tile_sums_kernel":
            return "exclusive_scan", True
        return (tgt or "exclusive_scan"), False
    return None, False


def main():
    kernels, listing = {}, []
    for path in sys.argv[1:]:
        state = {"reversals": 0}
        listing.append(f"# {path}")
        for fn, met in parse(path):
            nm, opens = bench_name(fn, state)
            t_us = met.get("gpu__time_duration.sum", 0.0) / 1e3
            listing.append(f"{t_us:10.1f} us  {int(met.get('smsp__inst_executed.sum', 0)):>12d} inst  "
                           f"{(met.get('dram__bytes_read.sum', 0) + met.get('dram__bytes_write.sum', 0)) / 1e6:9.1f} MB dram  "
                           f"issue {met.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):5.1f}%  "
                           f"L2 hit {met.get('lts__t_sector_hit_rate.pct', 0):5.1f}%  regs {int(met.get('launch__registers_per_thread', 0)):3d}  "
                           f"{nm or '-':<18s} {fn[:90]}")
            if nm is None:
                continue
            k = kernels.setdefault(nm, {"launches": 0, "inst": 0.0, "dram": 0.0, "ns": 0.0, "source_file": Path(path).name})
            if k["source_file"] != Path(path).name:
                continue  # the first capture that saw a kernel name wins (set construction appears in both)
            k["launches"] += 1 if opens else 0
            k["inst"] += met.get("smsp__inst_executed.sum", 0.0)
            k["dram"] += met.get("dram__bytes_read.sum", 0.0) + met.get("dram__bytes_write.sum", 0.0)
            k["ns"] += met.get("gpu__time_duration.sum", 0.0)
    out = {"source": "ncu --metrics smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                     "--clock-control none over one step of configs[1] (profiles/ncu_step.py): " + ", ".join(Path(p).name for p in sys.argv[1:]),
           "kernels": {}}
    total_ns = sum(k["ns"] for k in kernels.values())
    for nm, k in kernels.items():
        n = max(1, k["launches"])
        out["kernels"][nm] = {"launches_per_step": k["launches"], "warp_inst_per_launch": k["inst"] / n,
                              "dram_bytes_per_launch": k["dram"] / n, "ncu_us_per_launch": k["ns"] / n / 1e3,
                              "share_of_step_ncu": round(k["ns"] / total_ns, 4) if total_ns else 0.0, "capture": k["source_file"]}
    (Path(__file__).resolve().parent / "kernel_counters.json").write_text(json.dumps(out, indent=1) + "\n")
    print("\n".join(listing))


if __name__ == "__main__":
    main()
