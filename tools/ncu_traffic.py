#!/usr/bin/env python3
"""ncu --set full report -> profiles/ncu_traffic.json: DRAM bytes per FRAME per kernel (dram__bytes_read.sum +
dram__bytes_write.sum per launch / frames per launch, averaged over the launches of the kernel in the report).
usage: ncu_traffic.py report.ncu-rep frames_per_launch out.json "description of the capture" """
import csv
import json
import re
import subprocess
import sys

rep, fpl, out, desc = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]


def val(r, name):
    i = h.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


acc = {}
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    name = re.sub(r"\(.*$", "", name).replace("void ", "").strip()
    # the profiler facility of the library names template instantiations by their source spelling
    name = re.sub(r"^k_ccl_tile<.*>$", "k_ccl_tile<LinkFn>", name)
    name = re.sub(r"^kf_blb_stream4<.*>$", "kf_blb_stream4<npx>", name)
    name = re.sub(r"<\(bool\)1>$", "<true>", name)
    name = re.sub(r"<\(bool\)0>$", "<false>", name)
    name = re.sub(r"^(kr_reduceLS_list|kr_reduceLS)<\(int\)(\d)>$", r"\1<\2>", name)
    b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    t = float(r[h.index("gpu__time_duration.sum")].replace(",", ""))
    a = acc.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += b; a[2] += t
js = {"source": desc, "frames_per_launch": fpl,
      "kernels": {k: {"launches": v[0], "dram_bytes_per_frame": round(v[1] / v[0] / fpl), "ncu_time_per_launch": round(v[2] / v[0], 2)} for k, v in sorted(acc.items())}}
json.dump(js, open(out, "w"), indent=1)
print("wrote", out, len(acc), "kernels")
