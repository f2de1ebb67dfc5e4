#!/usr/bin/env python3
"""Per-CUDA-source-line share of executed warp instructions and stall samples for one kernel of an
ncu report captured with --import-source on (B200_PROFILING.md):

    python profiles/source_hotspots.py gpurun_out/prof_r1q.ncu-rep scan_spec [launch_skip] [top_n]
"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      f"regex:{kern}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
files, cur, hdr = {}, None, None
lines = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ii, si = r.index("Instructions Executed"), r.index("# Samples")
        continue
    if hdr is None or len(r) <= ii or not r[0].strip().isdigit():
        continue
    try:
        lines.append((int(r[ii]), int(r[si]), cur, int(r[0]), r[1].strip()[:100]))
    except ValueError:
        pass
ti, ts = sum(x[0] for x in lines), sum(x[1] for x in lines)
print(f"# {kern} (launch skip {skip}): {ti} warp instructions, {ts} stall samples")
print(f"{'inst%':>6} {'stall%':>6}  location")
for n, s, f, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{100 * n / ti:6.2f} {100 * s / max(ts, 1):6.2f}  {f}:{ln}  {src}")
