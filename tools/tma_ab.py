#!/usr/bin/env python3
"""A/B of the tensor-map (TMA) staging in kf_edge_thin against the plain staging: parity of the thinned edge strength against the CPU
oracle (production schedule stopped after stage 3) and the kernel's device time (CUDA events, 8 frames per launch), each mode in its
own process (RD_TMA is read once).   usage: tma_ab.py            (needs a GPU)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import math
    import numpy as np
    import torch
    import parity
    import rectdetect_b200 as rd
    from rectdetect_b200.synth import synth_frame
    dev = rd.Device(0)
    for iw, ih, seed in ((1280, 720, 2), (640, 480, 1), (1284, 724, 31), (96, 64, 31)):
        bad = [r for r in parity.compare_fast_stages(iw, ih, seed, [3], rd, dev) if r[2] != 0]
        print("  parity %dx%d seed %d (thinned edge strength vs oracle): %s" % (iw, ih, seed, "bit-exact" if not bad else bad))
    iw, ih, fpl = 1280, 720, 8
    frames = torch.empty((fpl, ih, 3 * iw), dtype=torch.uint8)
    for i in range(fpl):
        synth_frame(iw, ih, 1000 + i, out=frames[i].numpy())
    d = frames.cuda()
    b = rd.Batch(0, iw, ih, nctx=1, frames_per_launch=fpl)
    T = math.tan(math.radians(36))
    for _ in range(3):
        b.run(d.data_ptr(), ih * 3 * iw, 3 * iw, fpl, T, on_device=True, want_rects=False)
    ts = []
    for _ in range(5):
        rd.api.profile_start("kf_edge_thin")
        b.run(d.data_ptr(), ih * 3 * iw, 3 * iw, fpl, T, on_device=True, want_rects=False)
        prof = rd.api.profile_stop()
        ts += [v[1] / v[0] * 1e3 for v in prof.values()]
    print("  kernel %s: %.1f us per launch of %d frames 1280x720 (median of %d, CUDA events)" % (list(prof)[0], float(np.median(ts)), fpl, len(ts)))
    sys.exit(0)
for mode in ("0", "1"):
    print("RD_TMA=%s (%s)" % (mode, "tensor-map tile loads for interior CTAs" if mode == "1" else "plain staging"))
    sys.stdout.flush()
    subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, RD_TMA=mode))
