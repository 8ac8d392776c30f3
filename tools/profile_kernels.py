#!/usr/bin/env python3
"""Per-kernel device time of one frame through oclrect_executeOnce on a single stream (CUDA events around every launch,
rd_profile_start/stop).  usage: profile_kernels.py [iw ih [reps]]  -> table sorted by total time"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import math  # noqa: E402
import rectdetect_b200 as rd  # noqa: E402
from rectdetect_b200.synth import synth_frame  # noqa: E402

iw, ih = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1280, 720)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
T = math.tan(math.radians(36))
dev = rd.Device(0)
g = rd.OclRect(dev, iw, ih)
frames = [synth_frame(iw, ih, 1000 + i) for i in range(reps)]
for f in frames[:3]:
    g.execute_once(f, T)
t0 = time.perf_counter()
for f in frames:
    g.execute_once(f, T)
wall = (time.perf_counter() - t0) / reps
rd.api.profile_start(None)
for f in frames:
    g.execute_once(f, T)
prof = rd.api.profile_stop()
tot = sum(v[1] for v in prof.values())
print("%dx%d: wall %.3f ms/frame (unprofiled, single stream, incl. host tail); sum of kernel time %.3f ms/frame, %d launches/frame"
      % (iw, ih, wall * 1e3, tot / reps, sum(v[0] for v in prof.values()) // reps))
print("%-28s %8s %10s %10s %6s" % ("kernel", "n/frame", "us/launch", "us/frame", "%"))
for k, (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print("%-28s %8.1f %10.2f %10.1f %6.1f" % (k, c / reps, ms / c * 1e3, ms / reps * 1e3, 100 * ms / tot))
