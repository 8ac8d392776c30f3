#!/usr/bin/env python3
"""Per CUDA source line totals of one kernel in an ncu report captured with --import-source on (kernels built with -lineinfo):
warp-level instructions executed and stall samples, top lines first.
usage: ncu_lines.py report.ncu-rep kernel_regex [launch_index [top_n]]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
fname, func, h, out = "", "", None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) == 2 and r[0] == "Function Name":
        func = r[1]
    elif r and r[0] == "Line No":
        h = r
    elif h and len(r) == len(h) and r[0]:
        try:
            out.append((float(r[h.index("Instructions Executed")]), float(r[h.index("Warp Stall Sampling (All Samples)")]), fname, int(r[0]), r[1].strip()[:110]))
        except ValueError:
            pass
ti, ts = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
print(func[:120], "| warp instructions", int(ti), "| samples", int(ts))
for i, s, f, ln, src in sorted(out, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% stall  %s:%-4d %s" % (100 * i / ti, 100 * s / ts, f, ln, src))
