#!/usr/bin/env python3
"""Run on a GPU box: production-schedule parity (every stage plane + rectangle lists) against the CPU oracle over a sweep of
frame sizes and seeds, including widths that are not multiples of 4 / 32 / 128 and frames smaller than a tile.
Every third case is a frame tiled from 3x3 independent scenes (many more segments / regions / candidates: long read-back
records, long despeckle2 runs, more tail candidates).
usage: gpu_stress_parity.py [quick | <first seed> <seeds per size> [replay]]   (replay: labelMergeMain with the first pass replayed, both sides)"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import parity  # noqa: E402
import oracle_lib as ol  # noqa: E402
import rectdetect_b200 as rd  # noqa: E402

quick = len(sys.argv) == 2
sizes = [(640, 480), (641, 479), (322, 200), (130, 97), (96, 64), (1000, 562), (1284, 724), (1920, 1080), (257, 511), (48, 40), (1276, 716), (800, 600)]
seeds = [11, 12] if quick else [21, 22, 23, 24]
if len(sys.argv) > 2:
    seeds = list(range(int(sys.argv[1]), int(sys.argv[1]) + int(sys.argv[2])))
if "replay" in sys.argv[1:]:
    rd.set_merge_replay(True)
    ol.oracle().ora_set_merge_replay(1)
    print("labelMergeMain: first-pass replay on (CUDA and oracle)")
dev = rd.Device(0)
t0 = time.time()
nbad = 0
ncase = 0
for iw, ih in sizes:
    for seed in seeds:
        bad = [r for r in parity.compare_fast_stages(iw, ih, seed, sorted(parity.FAST_STAGES), rd, dev) if r[2] != 0]
        ncase += 1
        img = ol.dense_frame(iw, ih, seed, 3, 3) if (ncase % 3 == 0 and iw >= 320) else ol.synth_frame(iw, ih, seed)
        o = ol.OracleRect(iw, ih)
        g = rd.OclRect(dev, iw, ih)
        ok, why = True, ""
        for rep in range(2):          # the second pass sees the carried-over strength accumulator
            ra, rb = o.execute_once(img, parity.TAN_AOV), g.execute_once(img, parity.TAN_AOV)
            k, w = parity.rects_close(ra, rb)
            ok, why = ok and k, why + " " + w
        g.close()
        o.close()
        status = "OK  " if (not bad and ok) else "FAIL"
        nbad += status == "FAIL"
        print("%s %4dx%-4d seed %3d rects:%s %s" % (status, iw, ih, seed, why, bad[:3] if bad else ""), flush=True)
print("%d failures, %.0f s" % (nbad, time.time() - t0))
sys.exit(1 if nbad else 0)
