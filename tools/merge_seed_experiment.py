#!/usr/bin/env python3
"""Where does the order dependence of labelMergeMain live?  For every frame of the CPU sweeps: the reference's label plane after its
k-th raster pass (k = 0: labelxPreprocess only, 1, 2) is handed to the oracle's schedule-independent fixed point as its seed, and the
result is compared with the reference's plane after all 8 passes (oracle/_ref/librd_ref.so, work-items in raster order).
Outcome (profiles/r04t_merge_seed_experiment.txt): seeded with the plane after the FIRST pass the fixed point is the reference's own
label plane - image frame included - on 24 of 27 frames (3 / 1 / 87 pixels on the others), after the third pass on all - which
is why the CUDA path replays exactly that pass (rd_merge1.cuh).
usage: merge_seed_experiment.py [big]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import oracle_lib as ol  # noqa: E402
import test_ref_device as t  # noqa: E402

P = t.P
sizes = [(640, 480), (641, 479), (322, 200), (130, 97), (96, 64), (48, 40), (257, 511), (1000, 562), (1280, 720)]
cases = [(iw, ih, s) for iw, ih in sizes for s in (31, 32, 33)]
if "big" in sys.argv[1:]:
    cases = [(1920, 1080, 2000), (1920, 1080, 2001), (3840, 2160, 5)]
L = ol.oracle()
tot, same = [0, 0, 0, 0], [0, 0, 0, 0]
print("label pixels (whole plane) that differ from the reference's 8 raster passes; fixed point seeded with the reference's plane after pass k")
print("%-18s %12s %12s %12s %12s" % ("frame", "k=0 (default)", "k=1 (replay)", "k=2", "k=3"))
for iw, ih, seed in cases:
    L.ora_set_merge_replay(0)
    _, d = t._oracle_stage_b_inputs(iw, ih, seed)
    pre = np.zeros(iw * ih, np.int32)
    k_pre = t.rl.kernel_direct("rect", "labelxPreprocess")
    k_pre.argtypes = [t.ci, t.ci, t.vp, t.vp, t.ci, t.ci]
    k_pre(iw, ih, P(pre), P(d["pix"]), iw, ih)
    refs = t._ref_label_merge(d, iw, ih, passes=8)
    row = []
    for k, plane in enumerate([pre, refs[0], refs[1], refs[2]]):
        lab = plane.copy()
        L.ora_rect_labelMerge_seeded(P(lab), P(d["pix"]), P(d["mask"]), P(d["edge"]), iw, ih)
        off = int((lab != refs[7]).sum())
        row.append(off)
        tot[k] += off
        same[k] += off == 0
    print("%-18s %12d %12d %12d %12d" % ("%dx%d s%d" % (iw, ih, seed), *row), flush=True)
print("%d frames; identical to the reference: %s; pixels off in total: %s" % (len(cases), same, tot))
print("(k=0 here is the seeded fixed point WITHOUT the top-row rule of the default mode: image-frame labels stay preprocess pointers)")
