#!/usr/bin/env python3
"""CPU sweep: the oracle against THE REFERENCE ITSELF (oracle/_ref/librd_ref.so, raster-order schedule) over frame sizes
(incl. widths that are not multiples of 32, frames smaller than the blur / window radii), row strides wider than the image,
blank frames and many seeds.  Per case: which of the order-independent planes are bit-exact after the full genGPUTask, how far
the region map is off (labelMergeMain is order dependent: the one remaining canonical substitute), the vote table on identical
inputs, and how the rectangle lists compare.   usage: ref_vs_oracle_sweep.py [quick|big] [replay]   (needs a built librd_ref.so;
`replay`: the oracle with the first labelMergeMain pass replayed, ora_set_merge_replay(1))"""
import ctypes as C
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import oracle_lib as ol  # noqa: E402
import ref_lib as rl  # noqa: E402
from test_ref_device import _match_rects  # noqa: E402

TAN = math.tan(math.radians(36.0))
quick = "quick" in sys.argv[1:]
big = "big" in sys.argv[1:]
replay = "replay" in sys.argv[1:]          # configs 4 and 5 of BASELINE.json: 1920x1080 seed 2000, 3840x2160 seed 5
sizes = [(640, 480), (641, 479), (322, 200), (130, 97), (96, 64), (48, 40), (257, 511), (1000, 562), (1280, 720)]
seeds = [31] if quick else [31, 32, 33]
cases = [(iw, ih, s, None, False) for iw, ih in sizes for s in seeds]
cases += [(640, 360, 41, 4 * 640 - 3, False), (333, 217, 42, 3 * 333 + 5, False), (320, 240, 0, None, True)]
if big:
    cases = [(1920, 1080, 2000, None, False), (1920, 1080, 2001, None, False), (3840, 2160, 5, None, False)]
k_votes = rl.kernel_direct("rect", "reduceLS")
k_votes.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
rl.set_threads(1)
t0 = time.time()
bad = tot_r = tot_m = 0
print("%-26s %-58s %-12s %-7s %s" % ("case", "bit-exact planes (plab thin strong quant lsid ls)", "segid off", "votes", "rects ref/ora/matched"))
LO = ol.oracle()
LO.ora_set_merge_replay(1 if replay else 0)
print("oracle: labelMergeMain %s" % ("with the first pass replayed in raster order" if replay else "as the schedule-independent fixed point (default)"))


for iw, ih, seed, ws, blank in cases:
    n = iw * ih
    img = np.full((ih, 3 * iw), 128, np.uint8) if blank else ol.synth_frame(iw, ih, seed, ws=ws)
    stride = img.shape[-1]
    r = rl.RefRect(iw, ih)
    r.gpu_task(img, stride)
    a = r.cpu_task(TAN)
    o = ol.OracleRect(iw, ih)
    o.gpu_task(img, stride)
    b = o.cpu_task(TAN)
    flags = []
    for name in ("buf1", "buf3", "buf4", "buf0"):      # thinned strength, strong-edge bitmap, quantised colours, segment-id map
        flags.append(bool(np.array_equal(r.buffer(name), o.buffer(name))))
    flags.append(r.ls_list().tobytes() == o.ls_list().tobytes())
    r2 = rl.RefRect(iw, ih)
    r2.gpu_task(img, stride, 1)
    o2 = ol.OracleRect(iw, ih)
    o2.gpu_task(img, stride, 1)
    flags.insert(0, bool(np.array_equal(r2.buffer("buf0"), o2.buffer("buf0"))))
    r2.close()
    o2.close()
    seg_off = float((r.buffer("iobuf1") != o.buffer("iobuf1")).mean())
    out = np.zeros(4 * n, np.int32)
    seg, lsid = o.buffer("iobuf1").copy(), o.buffer("buf0").copy()
    k_votes(iw, ih, out.ctypes.data, seg.ctypes.data, lsid.ctypes.data, iw, ih, n * 4 // 5)
    votes_ok = bool(np.array_equal(out, o.buffer("ioBig1")))
    m = _match_rects(a, b)
    tot_r += max(len(a), len(b))
    tot_m += m
    ok = all(flags) and votes_ok
    bad += 0 if ok else 1
    print("%-26s %-58s %-12s %-7s %s" % ("%dx%d s%d%s%s" % (iw, ih, seed, " ws%d" % ws if ws else "", " blank" if blank else ""),
                                                     " ".join("yes" if f else "NO " for f in flags), "%.3f %%" % (100 * seg_off), "yes" if votes_ok else "NO",
                                                     "%d/%d/%d" % (len(a), len(b), m)))
    r.close()
    o.close()
print("%d cases, %d with a plane that is not bit-exact; rectangles: %d of %d matched within 1e-4; %.0f s" % (len(cases), bad, tot_m, tot_r, time.time() - t0))
