#!/usr/bin/env python3
"""Run on a GPU box: step-by-step parity report of the CUDA rect pipeline against the CPU oracle (not a test;
prints every mismatch instead of stopping at the first).  usage: gpu_parity_report.py [iw ih seed] [steps...]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import parity  # noqa: E402
import oracle_lib as ol  # noqa: E402
import rectdetect_b200 as rd  # noqa: E402

iw, ih, seed = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (640, 480, 1)
steps = [int(a) for a in sys.argv[4:]] or sorted(parity.STEP_BUFFERS)
print("devices:", rd.device_count(), rd.lib().rd_version())
dev = rd.Device(0)
t0 = time.time()
for k, name, bad, total, detail in parity.compare_steps(iw, ih, seed, steps, rd, dev):
    print("step %2d %-7s %s %d/%d %s" % (k, name, "OK  " if bad == 0 else "FAIL", bad, total, detail), flush=True)
print("steps took %.1f s" % (time.time() - t0))
for k, name, bad, total, detail in parity.compare_fast_stages(iw, ih, seed, sorted(parity.FAST_STAGES), rd, dev):
    print("fast stage %2d %-7s %s %d/%d %s" % (k, name, "OK  " if bad == 0 else "FAIL", bad, total, detail), flush=True)
# whole pipeline through the public entry point
img = ol.synth_frame(iw, ih, seed)
o = ol.OracleRect(iw, ih)
g = rd.OclRect(dev, iw, ih)
for rep in range(2):                      # second pass exercises the cross-frame carry-over (Q1)
    ra = o.execute_once(img, parity.TAN_AOV)
    t0 = time.time()
    rb = g.execute_once(img, parity.TAN_AOV)
    dt = time.time() - t0
    print("executeOnce pass %d: oracle %d rects, cuda %d rects, %s, %.2f ms, launches so far %d" % (rep, len(ra), len(rb), parity.rects_close(ra, rb)[1], dt * 1e3, rd.kernel_launches()))
g.close()
