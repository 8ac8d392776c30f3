#!/usr/bin/env python3
"""Condensed view of an ncu report: one block per kernel launch with the metrics that matter for this project.
usage: ncu_summary.py report.ncu-rep [max_launches]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
mx = int(sys.argv[2]) if len(sys.argv) > 2 else 100
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
short = {
    "gpu__time_duration.sum": "time_us", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram%",
    "dram__bytes_read.sum": "dram_rd_MB", "dram__bytes_write.sum": "dram_wr_MB", "sm__warps_active.avg.pct_of_peak_sustained_active": "occ%",
    "launch__registers_per_thread": "regs", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue%", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_conflicts",
    "smsp__inst_executed.sum": "warp_inst", "lts__t_sector_hit_rate.pct": "l2hit%", "launch__grid_size": "grid", "launch__block_size": "block",
}
stalls = [c for c in h if c.startswith("smsp__average_warps_issue_stalled_") and c.endswith("_per_issue_active.ratio")]
for r in rows[2:2 + mx]:
    name = r[h.index("Kernel Name")].split("(")[0]
    vals = []
    for k, s in short.items():
        if k in h:
            v = r[h.index(k)]
            try:
                v = "%.4g" % float(v.replace(",", ""))
            except ValueError:
                pass
            vals.append("%s=%s" % (s, v))
    st = sorted(((float(r[h.index(c)] or 0), c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for c in stalls), reverse=True)[:4]
    print(name, " ".join(vals))
    print("    stalls:", ", ".join("%s %.2f" % (n, v) for v, n in st))
