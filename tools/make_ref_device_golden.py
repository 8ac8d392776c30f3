#!/usr/bin/env python3
"""Mint tests/golden/ref_device_golden.json from THE REFERENCE ITSELF (oracle/_ref/librd_ref.so, built by
`make -C oracle _ref` from /root/reference: its unmodified host code + its OpenCL C kernels compiled as C++, work-items in
raster order).  For each configured frame the reference's genGPUTask runs and the SHA-256 of the planes that do not depend
on the order of its work-items are recorded: packed Lab, thinned edge strength (floats), edge bitmap #1, string labels,
strong-edge bitmap, blurred / quantised colours, segment-id map, and the polyline vertex list (LS_t records).
The CPU tests replay the oracle against these hashes, the GPU tests the CUDA pipeline.
usage: python tools/make_ref_device_golden.py   (needs /root/reference or a built oracle/_ref/librd_ref.so)
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import oracle_lib as ol  # noqa: E402
import ref_lib as rl  # noqa: E402

CASES = [(640, 480, 1), (333, 217, 7), (1280, 720, 2), (1280, 720, 1000)]
# plane -> (reference launch count to stop after, reference buffer, element count as a multiple of n or "ls")
PLANES = {
    "plab": (1, "buf0", 1), "thin_strength_f32": (30, "buf1", 1), "edge_bitmap1": (32, "tmp1", 1), "string_labels": (47, "buf2", 1),
    "blurred_plab": (71, "buf4", 1), "quantised_plab": (73, "buf4", 1), "strong_edge": (75, "buf3", 1),
    "lsid": (220, "buf0", 1), "ls": (220, "ioBig0", "ls"),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def reference_planes(iw, ih, seed):
    img = ol.synth_frame(iw, ih, seed)
    n = iw * ih
    out = {}
    for name, (limit, buf, count) in PLANES.items():
        r = rl.RefRect(iw, ih)
        r.gpu_task(img, img.shape[-1], limit)
        out[name] = r.ls_list().view(np.int32).copy() if count == "ls" else r.buffer(buf)[: n * count].copy()
        r.close()
    return out


def main():
    rl.set_threads(1)
    fixtures = []
    for iw, ih, seed in CASES:
        pl = reference_planes(iw, ih, seed)
        fixtures.append({"iw": iw, "ih": ih, "seed": seed, "n_ls": int(pl["ls"][0]), "sha": {k: sha(v) for k, v in pl.items()}})
        print(iw, ih, seed, "segments", int(pl["ls"][0]))
    path = os.path.join(ROOT, "tests", "golden", "ref_device_golden.json")
    json.dump(fixtures, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
