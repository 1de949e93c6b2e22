#!/usr/bin/env python
"""brief per-launch table from an .ncu-rep (raw page): time, DRAM bytes, pipes, occupancy, top stalls"""
import csv, subprocess, sys, re
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
H = {h: i for i, h in enumerate(hdr)}
def g(r, k, d=0.0):
    try: return float(r[H[k]].replace(",", ""))
    except Exception: return d
for r in rows[2:]:
    name = r[H["Kernel Name"]].split("(")[0]
    stalls = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): g(r, h) for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
    tot = sum(stalls.values()) or 1
    top = ", ".join(f"{k} {v/tot:.0%}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
    rd, wr = g(r, "dram__bytes_read.sum"), g(r, "dram__bytes_write.sum")
    ur, uw = r[1][H["dram__bytes_read.sum"]] if False else rows[1][H["dram__bytes_read.sum"]], rows[1][H["dram__bytes_write.sum"]]
    print(f"{name:28s} grid {r[H['launch__grid_size']]:>6s} t {g(r,'gpu__time_duration.sum'):8.1f}{rows[1][H['gpu__time_duration.sum']]} dram r {rd:.1f}{ur} w {wr:.1f}{uw} "
          f"regs {r[H['launch__registers_per_thread']]} warps {g(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f}% "
          f"fmaheavy {g(r,'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed'):.0f}% alu {g(r,'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed'):.0f}% "
          f"issue {g(r,'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f}% dram% {g(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f}\n      stalls: {top}")
