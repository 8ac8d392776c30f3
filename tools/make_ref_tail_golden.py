#!/usr/bin/env python3
"""Writes tests/golden/ref_tail_golden.json: rectangle lists produced by the REFERENCE's own host tail (executeCPUTask of
/root/reference/oclrect.c, compiled by `make -C oracle _ref`) from the oracle's device-stage outputs of seeded synthetic
frames.  Run in the build container (needs /root/reference); the fixture lets the tails be checked where the reference
sources are absent.  Floats are stored as hex strings: the comparison is bit-exact."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np  # noqa: E402
import oracle_lib as ol  # noqa: E402
import parity  # noqa: E402
import ref_tail_lib as rt  # noqa: E402

CASES = [(640, 480, 1), (640, 480, 2), (333, 217, 7), (640, 360, 1001), (1280, 720, 2)]
out = []
for iw, ih, seed in CASES:
    o = ol.OracleRect(iw, ih)
    o.gpu_task(ol.synth_frame(iw, ih, seed))
    r = rt.execute_cpu_task(o.buffer("ioBig0"), o.buffer("ioBig1"), o.buffer("iobuf1"), iw, ih, parity.TAN_AOV)
    out.append({"iw": iw, "ih": ih, "seed": seed, "n_rects": int(len(r)), "status": [int(s) for s in r["status"]],
                "c2": [float(v).hex() for v in r["c2"].reshape(-1)], "c3": [float(v).hex() for v in r["c3"].reshape(-1)],
                "value": [float(v).hex() for v in r["value"]]})
    o.close()
    print(iw, ih, seed, len(r), "rectangles")
path = os.path.join(ol.ROOT, "tests", "golden", "ref_tail_golden.json")
json.dump(out, open(path, "w"), indent=0)
print("wrote", path)
