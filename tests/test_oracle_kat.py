"""Known-answer tests that pin the CPU oracle (CPU only).

The reference has no tests, fixtures or golden vectors, so the oracle is pinned by closed forms,
independent re-implementations (numpy / scipy) and self-generated fixtures (tests/golden/, tools/make_golden.py).
"""
import ctypes as C
import json
import math
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol
from tools_path import ROOT

L = ol.oracle()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------ tables (oclimgutil.cl:661-1125)
def test_tables_match_pinned_digests():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_tables.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


# ------------------------------------------------------------------ srgb2plab (oclimgutil.cl:106-134) vs float CIE Lab
def _lab_float(b, g, r):
    def lin(c):
        c = c / 255.0
        return np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)
    R, G, B = lin(r), lin(g), lin(b)
    X = (0.412453 * R + 0.357580 * G + 0.180423 * B) / 0.950456
    Y = 0.212671 * R + 0.715160 * G + 0.072169 * B
    Z = (0.019334 * R + 0.119193 * G + 0.950227 * B) / 1.088754
    f = lambda t: np.where(t > 0.008856, np.cbrt(t), 7.787 * t + 16.0 / 116.0)
    Lc = np.where(Y > 0.008856, 116.0 * np.cbrt(Y) - 16.0, 903.3 * Y)
    return Lc, 500.0 * (f(X) - f(Y)), 200.0 * (f(Y) - f(Z))


def test_srgb2plab_tracks_float_lab():
    rng = np.random.default_rng(0)
    cols = rng.integers(0, 256, size=(4000, 3))
    cols = np.vstack([cols, [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [128, 128, 128]]])
    for b, g, r in cols:
        v = L.ora_srgb2plab(int(b), int(g), int(r))
        l, a, bb = v & 4095, (v >> 12) & 1023, (v >> 22) & 1023
        Lf, af, bf = _lab_float(float(b), float(g), float(r))
        # packing: l/4096 = L*/255, a/1024 = (a* + 128)/256, b/1024 = (b* + 128)/256 (half-LSB offsets aside)
        assert abs(l * 255.0 / 4096.0 - Lf) < 0.3, (b, g, r, l, Lf)
        assert abs((a - 514) / 4.0 - af) < 1.0, (b, g, r, a, af)
        assert abs((bb - 514) / 4.0 - bf) < 1.0, (b, g, r, bb, bf)


def test_srgb2plab_grey_axis_is_neutral_and_monotone():
    prev = -1
    for g in range(256):
        v = L.ora_srgb2plab(g, g, g)
        l, a, b = v & 4095, (v >> 12) & 1023, (v >> 22) & 1023
        assert abs(a - 514) <= 2 and abs(b - 514) <= 2
        assert l >= prev
        prev = l
    assert L.ora_srgb2plab(0, 0, 0) & 4095 == 0
    assert abs((L.ora_srgb2plab(255, 255, 255) & 4095) - 100.0 * 4096 / 255) < 2


# ------------------------------------------------------------------ pack / unpack (oclimgutil.cl:28-39)
def test_pack_unpack_roundtrip_all_codes():
    out = np.zeros(3, np.float32)
    rng = np.random.default_rng(1)
    for v in list(rng.integers(0, 2 ** 32, size=3000, dtype=np.uint64)) + [0, 2 ** 32 - 1]:
        v = int(v)
        L.ora_unpacklab(v, _p(out))
        assert L.ora_packlab(float(out[0]), float(out[1]), float(out[2])) == v


def test_packlab_saturates_negative_and_large():
    assert L.ora_packlab(-1.0, -0.5, -3.0) == 0
    assert L.ora_packlab(2.0, 2.0, 2.0) == (1023 << 22) | (1023 << 12) | 4095
    assert L.ora_packlab(float("nan"), 0.0, 0.0) & 4095 == 0


# ------------------------------------------------------------------ border rules (oclimgutil.cl:41-63)
def test_mirror_and_repeat():
    iw = 10
    assert [L.ora_mirror1(x, iw) for x in range(-3, 13)] == [3, 2, 1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 8, 7, 6]
    assert [L.ora_repeat1(x, iw) for x in (-3, -1, 0, 9, 10, 12)] == [7, 9, 0, 9, 0, 2]


# ------------------------------------------------------------------ xrandom (oclimgutil.cl:182-193) - independent python restatement
def _xrandom_py(s):
    M = (1 << 64) - 1
    def rotl(t, n):
        n &= 63
        return t if n == 0 else ((t << n) | (t >> (64 - n))) & M
    t = s
    for sh, k in ((24, 0xf3dd0fb7820fde37), (6, 0xe6c6ac2c59e52811), (18, 0x2fc7871fff7c5b45), (48, 0x47c7e1f70aa4f7c5),
                  (0, 0x094f02b7fb9ba895), (12, 0x89afda817e744570), (36, 0xc7277d052c7bf14b)):
        t = rotl(t, (s >> sh) & 63) ^ k
    return t


def test_xrandom_against_python():
    rng = np.random.default_rng(2)
    for s in [0, 1, 63, 64, 2 ** 63, 2 ** 64 - 1] + [int(v) for v in rng.integers(0, 2 ** 63, size=200, dtype=np.uint64)]:
        assert L.ora_xrandom(s) == _xrandom_py(s)
    M = (1 << 64) - 1
    for x in (0, 1, 12345, 921599):
        s = (((x ^ 0xb21c2cb635b48285) * 0x9b923b9cec745401) + ((0 ^ 0x7bb93d75a79d2f15) * 0x22cab58ada573a29)) & M
        want = _xrandom_py(s) & 0xffffffff
        assert (L.ora_rand_at(x, 0) & 0xffffffff) == want


# ------------------------------------------------------------------ recursive Gaussian (oclimgutil.cl:542-637)
def _blur(img, r=2):
    ih, iw = img.shape
    src = np.ascontiguousarray(img, np.float32)
    out, t0, t1 = np.zeros_like(src), np.zeros_like(src), np.zeros_like(src)
    L.ora_iirblur_f_f(_p(out), _p(src), _p(t0), _p(t1), r, iw, ih)
    return out


def test_iirblur_dc_gain_and_symmetry():
    c = _blur(np.full((48, 64), 0.37, np.float32))
    assert np.allclose(c, 0.37, atol=2e-5)
    imp = np.zeros((65, 65), np.float32)
    imp[32, 32] = 1.0
    k = _blur(imp)
    assert abs(float(k.sum()) - 1.0) < 1e-3
    assert np.allclose(k, k[::-1, :], atol=1e-6) and np.allclose(k, k[:, ::-1], atol=1e-6) and np.allclose(k, k.T, atol=1e-6)
    assert k[32, 32] == k.max() and k[32, 32] > k[32, 33] > k[32, 34] > k[32, 36] > 0


def test_iirblur_is_linear_within_rounding():
    rng = np.random.default_rng(3)
    a, b = rng.random((40, 56), np.float32), rng.random((40, 56), np.float32)
    assert np.allclose(_blur(a) + _blur(b), _blur(a + b), atol=5e-6)


# ------------------------------------------------------------------ label8x (oclimgutil.cl:495-538) vs scipy
def test_label8x_matches_scipy_components_and_reference_fixed_point():
    from scipy import ndimage
    rng = np.random.default_rng(4)
    for shape, p in (((37, 53), 0.45), ((64, 64), 0.6), ((50, 80), 0.3)):
        pix = (rng.random(shape) < p).astype(np.int32)
        ih, iw = shape
        out, tmp = np.zeros(shape, np.int32), np.zeros(shape, np.int32)
        passes = L.ora_label8x_int_int(_p(out), _p(pix), _p(tmp), 0, iw, ih)
        assert passes > 0, "union-find labels differ from the fixed point of the reference kernel"
        lab, n = ndimage.label(pix, structure=np.ones((3, 3)))
        assert (out[pix == 0] == -1).all()
        for c in range(1, n + 1):
            idx = np.flatnonzero(lab.reshape(-1) == c)
            assert (out.reshape(-1)[idx] == idx.min()).all()
        # bgc = -1: both values are labelled
        passes = L.ora_label8x_int_int(_p(out), _p(pix), _p(tmp), -1, iw, ih)
        assert passes > 0 and (out >= 0).all()


# ------------------------------------------------------------------ host-tail geometry (oclrect.c:418, 744-802)
def test_clip_line_cases():
    out = np.zeros(4)
    L.ora_clip_line(-5.0, 5.0, 15.0, 5.0, 0.0, 0.0, 10.0, 10.0, _p(out))
    assert out.tolist() == [0.0, 5.0, 10.0, 5.0]
    L.ora_clip_line(2.0, 2.0, 8.0, 9.0, 0.0, 0.0, 10.0, 10.0, _p(out))
    assert out.tolist() == [2.0, 2.0, 8.0, 9.0]
    L.ora_clip_line(-5.0, -5.0, -1.0, 20.0, 0.0, 0.0, 10.0, 10.0, _p(out))
    assert np.isnan(out).all()
    L.ora_clip_line(-10.0, -10.0, 20.0, 20.0, 0.0, 0.0, 10.0, 10.0, _p(out))
    assert np.allclose(out, [0, 0, 10, 10])


def test_intersection2():
    u, v, out = np.array([0.0, 0.0, 10.0, 0.0]), np.array([5.0, -5.0, 5.0, 5.0]), np.zeros(2)
    L.ora_intersection2(_p(u), _p(v), _p(out))
    assert np.allclose(out, [5.0, 0.0])
    v = np.array([0.0, 1.0, 10.0, 1.0])
    L.ora_intersection2(_p(u), _p(v), _p(out))
    assert np.isnan(out).all()


def test_pose_recovers_a_projected_rectangle():
    iw, ih, tan_aov = 1280, 720, math.tan(math.radians(36.0))
    f = (iw // 2) / tan_aov
    w, h, depth, yaw = 1.6, 0.9, 4.0, math.radians(25)
    pts3 = []
    for sx, sy in ((-1, 1), (1, 1), (1, -1), (-1, -1)):
        x, y, z = sx * w / 2, sy * h / 2, 0.0
        x, z = x * math.cos(yaw) + z * math.sin(yaw), -x * math.sin(yaw) + z * math.cos(yaw)
        pts3.append((x + 0.3, y - 0.1, z + depth))
    c2 = np.array([[f * X / Z + iw // 2, -f * Y / Z + ih // 2] for X, Y, Z in pts3])
    out = np.zeros(1, ol.RECT_DTYPE)
    L.ora_pose(_p(c2), iw, ih, tan_aov, _p(out))
    r = out[0]
    assert r["value"] < 1e-3
    c3 = r["c3"]
    e = [np.linalg.norm(c3[i] - c3[(i + 1) % 4]) for i in range(4)]
    aspect = max(e[0], e[1]) / min(e[0], e[1])
    assert abs(aspect - w / h) < 0.03
    assert r["status"] == 1


# ------------------------------------------------------------------ whole-pipeline fixtures (tests/golden/)
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_golden.json")))


@pytest.mark.parametrize("g", GOLDEN, ids=lambda g: "%dx%d-s%d" % (g["iw"], g["ih"], g["seed"]))
def test_oracle_pipeline_matches_golden(g):
    from tools_path import load_make_golden
    d = load_make_golden().fixture(g["iw"], g["ih"], g["seed"])
    for k in g:
        if k == "rect_c2":
            assert np.allclose(np.array(d[k]), np.array(g[k]), rtol=1e-9, atol=1e-9)
        else:
            assert d[k] == g[k], k


# ------------------------------------------------------------------ fixtures generated by THE REFERENCE (tools/make_ref_device_golden.py)
REF_DEVICE_GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_device_golden.json")))
# plane -> (oracle stop step, oracle buffer)
REF_DEVICE_PLANES = {"plab": (1, "buf0"), "thin_strength_f32": (7, "buf1"), "edge_bitmap1": (8, "tmp1"), "string_labels": (10, "buf2"),
                     "blurred_plab": (13, "buf4"), "quantised_plab": (14, "buf4"), "strong_edge": (15, "buf3"), "lsid": (20, "buf0"), "ls": (20, "ioBig0")}


@pytest.mark.parametrize("g", REF_DEVICE_GOLDEN, ids=lambda g: "%dx%d-s%d" % (g["iw"], g["ih"], g["seed"]))
def test_oracle_matches_reference_generated_planes(g):
    """SHA-256 of planes the reference's own kernels produced here (oracle/_ref/librd_ref.so): packed Lab, thinned edge strength
    (floats), edge bitmap, string labels, blurred / quantised colours, strong-edge bitmap, segment-id map, polyline vertex list"""
    import hashlib
    iw, ih = g["iw"], g["ih"]
    img = ol.synth_frame(iw, ih, g["seed"])
    for name, (step, buf) in REF_DEVICE_PLANES.items():
        o = ol.OracleRect(iw, ih)
        o.gpu_task(img, stop_step=step)
        a = o.ls_list().view(np.int32) if name == "ls" else o.buffer(buf)[: iw * ih]
        assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() == g["sha"][name], name
        o.close()


def test_oracle_detects_the_planted_quads():
    iw, ih, seed = 1280, 720, 2
    img, quads = ol.synth_frame(iw, ih, seed, with_truth=True)
    ol.oracle().ora_reset_stats()                    # the counters are process-wide
    o = ol.OracleRect(iw, ih)
    rects = o.execute_once(img, math.tan(math.radians(36.0)))
    screens = rects[(rects["status"] & 1) == 1]
    hits = 0
    for q in quads:
        q = q.reshape(4, 2)
        for r in screens:
            d = np.linalg.norm(r["c2"][:, None, :] - q[None, :, :], axis=2)
            if (d.min(axis=1) < 4.0).all():
                hits += 1
                break
    # later quads are painted over earlier ones, so only the unoccluded ones can come back with all four corners
    assert hits >= len(quads) // 3, "only %d of %d planted quadrilaterals were found" % (hits, len(quads))
    st = ol.stats()
    assert st["ls_overflow"] == 0 and st["mkpl_ties"] == 0


def test_nv12_to_bgr_equals_opencv_golden():
    """the NV12 front end against OpenCV-generated vectors (tools/make_nv12_golden.py), and against cv2 itself where
    it can be imported"""
    import hashlib
    import json
    import os
    golden = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nv12_golden.json")))
    for g in golden:
        rng = np.random.default_rng(g["seed"])
        iw, ih = g["iw"], g["ih"]
        yuv = rng.integers(0, 256, (ih * 3 // 2, iw), dtype=np.uint8)
        if g["corners"]:
            yuv[:] = rng.choice(np.array([0, 15, 16, 17, 128, 234, 235, 236, 255], np.uint8), yuv.shape)
        bgr = ol.nv12_to_bgr(yuv, iw, ih)
        assert [int(v) for v in bgr.reshape(-1)[:24]] == g["bgr_head"]
        assert hashlib.sha256(bgr.tobytes()).hexdigest() == g["bgr_sha"]
        try:
            import cv2
        except ImportError:
            continue
        assert np.array_equal(bgr.reshape(ih, iw, 3), cv2.cvtColor(yuv, cv2.COLOR_YUV2BGR_NV12))
    # a row stride wider than the frame
    yuv = np.random.default_rng(9).integers(0, 256, (48 * 3 // 2, 80), dtype=np.uint8)
    a = ol.nv12_to_bgr(yuv, 64, 48)
    b = ol.nv12_to_bgr(np.ascontiguousarray(yuv[:, :64]), 64, 48)
    assert np.array_equal(a, b)
