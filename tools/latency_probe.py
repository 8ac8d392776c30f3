#!/usr/bin/env python3
"""Latency of the reference's own API on one device: oclrect_executeOnce and the enqueue/poll stream (vidrect.cpp:159-172), per frame
size, plus the per-kernel device time of one frame (CUDA events, library profiler).   usage: latency_probe.py [w h]..."""
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import rectdetect_b200 as rd  # noqa: E402
from rectdetect_b200.synth import synth_frame  # noqa: E402

TAN = math.tan(math.radians(36.0))
sizes = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)] or [(1280, 720), (1920, 1080), (3840, 2160)]
dev = rd.Device(0)
for iw, ih in sizes:
    frames = [synth_frame(iw, ih, 1000 + i) for i in range(12)]
    g = rd.OclRect(dev, iw, ih)
    for f in frames[:3]:
        g.execute_once(f, TAN)
    once = []
    for f in frames:
        t0 = time.perf_counter()
        g.execute_once(f, TAN)
        once.append((time.perf_counter() - t0) * 1e3)
    t0 = time.perf_counter()
    g.enqueue_task(frames[0])
    for f in frames[1:]:
        g.enqueue_task(f)
        g.poll_task(TAN)
    g.poll_task(TAN)
    stream = (time.perf_counter() - t0) * 1e3 / len(frames)
    rd.api.profile_start(None, stages=True)
    g.run_device(frames[0], stop_step=0)
    prof = rd.api.profile_stop()
    ksum = sum(v[1] for v in prof.values())
    top = sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]
    print("%dx%d executeOnce median %.3f ms (min %.3f), stream %.3f ms/frame, kernel sum %.3f ms in %d launches; top: %s"
          % (iw, ih, float(np.median(once)), min(once), stream, ksum, sum(v[0] for v in prof.values()), ", ".join("%s %.0f us" % (k, v[1] * 1e3) for k, v in top)))
    g.close()
dev.close()
