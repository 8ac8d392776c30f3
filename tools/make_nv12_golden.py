#!/usr/bin/env python3
"""Golden vectors for the NV12 input path from OpenCV itself (cv2.cvtColor(..., COLOR_YUV2BGR_NV12), the conversion cv::VideoCapture
applies in front of the reference's programs): seeded random NV12 frames - every (Y, U, V) byte combination region gets hit, incl.
the saturating ones - and the SHA-256 of the BGR frames OpenCV produces, plus the first bytes for a readable diff.
usage: python tools/make_nv12_golden.py     (needs cv2; rewrites tests/golden/nv12_golden.json)"""
import hashlib
import json
import os

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = []
for iw, ih, seed in ((64, 48, 1), (130, 98, 2), (16, 16, 3)):
    rng = np.random.default_rng(seed)
    yuv = rng.integers(0, 256, (ih * 3 // 2, iw), dtype=np.uint8)
    if seed == 3:                                   # the corners of the code space
        yuv[:] = rng.choice(np.array([0, 15, 16, 17, 128, 234, 235, 236, 255], np.uint8), yuv.shape)
    bgr = cv2.cvtColor(yuv, cv2.COLOR_YUV2BGR_NV12)
    out.append({"iw": iw, "ih": ih, "seed": seed, "corners": seed == 3, "bgr_sha": hashlib.sha256(np.ascontiguousarray(bgr).tobytes()).hexdigest(),
                "bgr_head": [int(v) for v in bgr.reshape(-1)[:24]], "opencv": cv2.__version__})
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "nv12_golden.json"), "w"), indent=1)
print("wrote", len(out), "cases, OpenCV", cv2.__version__)
