#!/usr/bin/env python3
"""Per-kernel device time of the batched pipeline on ONE stream (nctx=1), frames resident on the device, no host tail:
CUDA events around every launch (rd_profile_start/stop).  usage: profile_batch.py [iw ih [fpl [nframes]]]"""
import math
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import rectdetect_b200 as rd  # noqa: E402
from rectdetect_b200.synth import synth_frame  # noqa: E402

iw, ih = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1280, 720)
fpl = int(sys.argv[3]) if len(sys.argv) > 3 else 16
nf = int(sys.argv[4]) if len(sys.argv) > 4 else 2 * fpl
T = math.tan(math.radians(36))
frames = torch.empty((nf, ih, 3 * iw), dtype=torch.uint8)
for i in range(nf):
    synth_frame(iw, ih, 1000 + i, out=frames[i].numpy())
d = frames.cuda()
b = rd.Batch(0, iw, ih, nctx=1, frames_per_launch=fpl)
for _ in range(2):
    b.run(d.data_ptr(), ih * 3 * iw, 3 * iw, nf, T, on_device=True, want_rects=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
b.run(d.data_ptr(), ih * 3 * iw, 3 * iw, nf, T, on_device=True, want_rects=False)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / nf
rd.api.profile_start(None)
b.run(d.data_ptr(), ih * 3 * iw, 3 * iw, nf, T, on_device=True, want_rects=False)
prof = rd.api.profile_stop()
tot = sum(v[1] for v in prof.values())
n = iw * ih
print("%dx%d, %d frames per launch, 1 stream, device-resident frames, no host tail: wall %.1f us/frame (unprofiled); sum of kernel time %.1f us/frame; %d launches per chunk"
      % (iw, ih, fpl, wall * 1e6, tot / nf * 1e3, sum(v[0] for v in prof.values()) * fpl // nf))
print("%-28s %8s %10s %10s %6s %9s" % ("kernel", "n/chunk", "us/launch", "us/frame", "%", "Mpx/s/1k"))
for k, (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print("%-28s %8.1f %10.2f %10.2f %6.1f %9.1f" % (k, c * fpl / nf, ms / c * 1e3, ms / nf * 1e3, 100 * ms / tot, n * fpl / (ms / c * 1e-3) / 1e9))
