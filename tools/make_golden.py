#!/usr/bin/env python3
"""Mint the golden fixtures of tests/golden/ from the CPU oracle.

The reference ships no tests or golden vectors and cannot run in this image, so these fixtures
pin the ORACLE against regressions; they are not outputs of the reference binary ("parity unpinned").
For each configured frame: SHA-256 of the edge bitmap, strong-edge mask, region (segid) map, segment-id map, the live
part of the line-segment list and the rectangle list, plus the small lists themselves.
usage: python tools/make_golden.py   (rewrites tests/golden/oracle_golden.json)
"""
import hashlib
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import oracle_lib as ol  # noqa: E402

CASES = [(640, 480, 1), (1280, 720, 2), (333, 217, 7), (1280, 720, 1000)]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def fixture(iw, ih, seed):
    img = ol.synth_frame(iw, ih, seed)
    o = ol.OracleRect(iw, ih)
    o.gpu_task(img, stop_step=8)
    edge1 = o.buffer("tmp1").copy()
    o.close()
    o = ol.OracleRect(iw, ih)
    rects = o.execute_once(img, math.tan(math.radians(36.0)))
    ls = o.ls_list()
    d = {
        "iw": iw, "ih": ih, "seed": seed, "frame_sha": sha(img),
        "edge_bitmap_sha": sha(edge1), "strong_edge_sha": sha(o.buffer("buf3")), "segid_sha": sha(o.buffer("iobuf1")),
        "lsid_sha": sha(o.buffer("buf0")), "ls_sha": sha(ls), "votes_sha": sha(o.buffer("ioBig1")), "rects_sha": sha(rects),
        "n_edge": int(edge1.sum()), "n_ls": int(len(ls) - 1), "n_rects": int(len(rects)),
        "rect_status": [int(s) for s in rects["status"]],
        "rect_c2": [[[float(v) for v in c] for c in r] for r in rects["c2"]],
    }
    o.close()
    return d


if __name__ == "__main__":
    out = [fixture(*c) for c in CASES]
    with open(os.path.join(ROOT, "tests", "golden", "oracle_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    for d in out:
        print(d["iw"], d["ih"], d["seed"], "edges", d["n_edge"], "ls", d["n_ls"], "rects", d["n_rects"])
