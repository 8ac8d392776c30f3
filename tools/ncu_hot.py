#!/usr/bin/env python3
"""Hot spots of one kernel in an ncu report: stall-reason totals, opcode mix and the top SASS lines by stall samples.
usage: ncu_hot.py report.ncu-rep kernel_regex [launch_index]"""
import collections
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[1]
cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
src, si, ie = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
tot, byop, execs, lines = collections.Counter(), collections.Counter(), collections.Counter(), []
for r in rows[2:]:
    if len(r) < len(h):
        continue
    for c in cols:
        try:
            tot[c] += float(r[h.index(c)])
        except ValueError:
            pass
    t = r[src].strip().split()
    op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")).split(".")[0]
    try:
        byop[op] += float(r[si])
        execs[op] += float(r[ie])
        lines.append((float(r[si]), r[src].strip()[:90], r[ie]))
    except ValueError:
        pass
n = sum(v for v, _, _ in lines) or 1
print(rows[0][1][:100], "| samples", int(n), "| sass lines", len(lines))
print("stalls:", ", ".join("%s %.0f%%" % (k[6:], 100 * v / n) for k, v in tot.most_common(7)))
print("by opcode (samples%, executed):", ", ".join("%s %.0f%% %d" % (k, 100 * v / n, execs[k]) for k, v in byop.most_common(10)))
for v, sx, e in sorted(lines, reverse=True)[:int(sys.argv[4]) if len(sys.argv) > 4 else 10]:
    print("  %5.1f%%  exec=%-9s %s" % (100 * v / n, e, sx))
