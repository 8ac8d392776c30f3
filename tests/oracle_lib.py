"""ctypes bindings for the CPU oracle (oracle/librd_oracle.so) and the synthetic frame generator.

Test infrastructure: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs only.  The product package (rectdetect_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "librd_oracle.so")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LS_DTYPE = np.dtype([("x0", "<f4"), ("y0", "<f4"), ("x1", "<f4"), ("y1", "<f4"),
                     ("startIndex", "<i4"), ("endIndex", "<i4"), ("leftPtr", "<i4"), ("rightPtr", "<i4"),
                     ("startCount", "<i4"), ("endCount", "<i4"), ("maxDist", "<i4"), ("polyid", "<i4"),
                     ("npix", "<i4"), ("level", "<i4")])
assert LS_DTYPE.itemsize == 56

RECT_DTYPE = np.dtype([("c2", "<f8", (4, 2)), ("c3", "<f8", (4, 3)), ("value", "<f8"), ("status", "<u4"), ("_pad", "<u4")])
assert RECT_DTYPE.itemsize == 176


class OraStats(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("label8x_seq_passes", "label8x_calls", "labelpl_components", "mkpl_ties",
                                       "mkpl_iterations_live", "vote_collisions", "vote_slots", "n_ls", "ls_overflow")]


def build_oracle(force=False):
    if force or not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return ORACLE_SO


_ora = None


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def oracle():
    global _ora
    if _ora is None:
        L = C.CDLL(build_oracle())
        vp, i, f, d = C.c_void_p, C.c_int, C.c_float, C.c_double
        sig = {
            "ora_set_threads": (None, [i]), "ora_get_threads": (i, []),
            "ora_get_stats": (None, [C.POINTER(OraStats)]), "ora_reset_stats": (None, []),
            "ora_clear": (None, [vp, i]), "ora_copy": (None, [vp, vp, i]),
            "ora_cast_i_f": (None, [vp, vp, f, i]), "ora_cast_c_i": (None, [vp, vp, i]),
            "ora_threshold_i_i": (None, [vp, vp, i, i, i, i]), "ora_threshold_f_f": (None, [vp, vp, f, f, f, i]),
            "ora_convert_plab_bgr": (None, [vp, vp, i, i, i]),
            "ora_unpack_f_f_f_plab": (None, [vp, vp, vp, vp, i, i]), "ora_pack_plab_f_f_f": (None, [vp, vp, vp, vp, i, i]),
            "ora_iirblur_f_f": (None, [vp, vp, vp, vp, i, i, i]),
            "ora_edgevec_f2_f": (None, [vp, vp, i, i]), "ora_edge_f_plab": (None, [vp, vp, i, i]),
            "ora_thinthres_f_f_f2": (None, [vp, vp, vp, i, i]), "ora_thincubic_f_f_f2": (None, [vp, vp, vp, i, i]),
            "ora_edgevec_f2_plab": (None, [vp, vp, i, i]), "ora_edge_f_f": (None, [vp, vp, i, i]),
            "ora_convert_bgr_plab": (None, [vp, vp, i, i, i]), "ora_convert_bgr_lumaf": (None, [vp, vp, f, i, i, i]),
            "ora_convert_bgr_labeli": (None, [vp, vp, i, i, i, i]), "ora_nv12_to_bgr": (None, [vp, vp, i, i, i, i]),
            "ora_label8x_int_int": (i, [vp, vp, vp, i, i, i]),
            "ora_calcStrength": (None, [vp, vp, vp, i, i]), "ora_filterStrength": (None, [vp, vp, i, i, i]),
            "ora_rect_simpleJunction": (None, [vp, vp, i, i]), "ora_rect_simpleConnect": (None, [vp, vp, i, i]),
            "ora_rect_stringify": (None, [vp, vp, i, i, i]),
            "ora_rect_blblur0": (None, [vp, vp, vp, i, i]), "ora_rect_blblur1": (None, [vp, vp, vp, i, i]),
            "ora_rect_quantize": (None, [vp, vp, i, i, i, i, i]), "ora_rect_despeckle": (None, [vp, vp, vp, i, i]),
            "ora_rect_mkMergeMask0": (None, [vp, vp, i, i]), "ora_rect_mkMergeMask1": (None, [vp, vp, i, i]),
            "ora_rect_labelMerge": (None, [vp, vp, vp, vp, i, i]),
            "ora_rect_labelMerge_first_pass": (None, [vp, vp, vp, vp, i, i]),
            "ora_set_merge_replay": (None, [i]),
            "ora_rect_labelMerge_seeded": (None, [vp, vp, vp, vp, i, i]),
            "ora_get_merge_replay": (i, []),
            "ora_rect_calcSize": (None, [vp, vp, i, i]), "ora_rect_despeckle2": (None, [vp, vp, i, i, i]),
            "ora_rect_markBoundary": (None, [vp, vp, i, i]), "ora_rect_reduceLS": (None, [vp, vp, vp, i, i, i]),
            "ora_polyline_execute": (None, [vp, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, f, i, i, i, i]),
            "ora_rect_create": (vp, [i, i]), "ora_rect_destroy": (None, [vp]),
            "ora_rect_gpu_task": (None, [vp, vp, i, i]), "ora_rect_buffer": (vp, [vp, C.c_char_p]),
            "ora_rect_cpu_task": (vp, [vp, d]), "ora_rect_execute_once": (vp, [vp, vp, i, d]),
            "ora_tail": (vp, [vp, vp, vp, i, i, d]), "ora_free": (None, [vp]),
            "ora_rect_last_times": (None, [vp, vp]),
            "ora_poly_frame": (None, [vp, i, i, i, f, i, i, vp, vp, vp]),
            "ora_srgb2plab": (C.c_uint32, [i, i, i]), "ora_packlab": (C.c_uint32, [f, f, f]), "ora_unpacklab": (None, [C.c_uint32, vp]),
            "ora_mirror1": (i, [i, i]), "ora_repeat1": (i, [i, i]),
            "ora_xrandom": (C.c_uint64, [C.c_uint64]), "ora_rand_at": (C.c_int32, [i, C.c_uint64]),
            "ora_clip_line": (None, [d, d, d, d, d, d, d, d, vp]), "ora_intersection2": (None, [vp, vp, vp]),
            "ora_pose": (None, [vp, i, i, d, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _ora = L
    return _ora


from rectdetect_b200.synth import synth_frame, synth_lib  # noqa: E402,F401  (the generator is workload input, not oracle)


def dense_frame(iw, ih, seed, tiles_x=4, tiles_y=4):
    """a frame tiled from tiles_x x tiles_y independent synthetic frames: many more rectangles / line segments than
    synth_frame gives (used to exercise the large read-back records; 1280x720 in 5x5 tiles has > 400 segments)"""
    tw, th = iw // tiles_x, ih // tiles_y
    img = np.full((ih, 3 * iw), 128, np.uint8)
    for ty in range(tiles_y):
        for tx in range(tiles_x):
            t = synth_frame(tw, th, seed * 100 + ty * tiles_x + tx)
            img[ty * th:(ty + 1) * th, 3 * tx * tw:3 * (tx + 1) * tw] = t[:, :3 * tw]
    return img


def bgr_to_nv12(img, iw, ih, ys=None):
    """a plausible NV12 rendition of a BGR frame (BT.601 limited range, 2x2 chroma average): workload input for the NV12 path; any byte
    pattern is a valid NV12 frame, so nothing depends on how faithful this is.  -> (ih * 3 // 2, ys) uint8"""
    ys = ys or iw
    bgr = np.ascontiguousarray(img)[:, : 3 * iw].reshape(ih, iw, 3).astype(np.float32)
    b, g, r = bgr[..., 0], bgr[..., 1], bgr[..., 2]
    y = 16 + 0.257 * r + 0.504 * g + 0.098 * b
    u = 128 - 0.148 * r - 0.291 * g + 0.439 * b
    v = 128 + 0.439 * r - 0.368 * g - 0.071 * b
    out = np.zeros((ih * 3 // 2, ys), np.uint8)
    out[:ih, :iw] = np.clip(np.rint(y), 0, 255)
    uv = np.stack([u.reshape(ih // 2, 2, iw // 2, 2).mean((1, 3)), v.reshape(ih // 2, 2, iw // 2, 2).mean((1, 3))], -1)
    out[ih:, :iw] = np.clip(np.rint(uv), 0, 255).reshape(ih // 2, iw)
    return out


def nv12_to_bgr(nv12, iw, ih, ws=None):
    """the oracle's NV12 -> BGR (= OpenCV's COLOR_YUV2BGR_NV12) -> (ih, ws) uint8"""
    ws = ws or 3 * iw
    nv12 = np.ascontiguousarray(nv12)
    out = np.zeros((ih, ws), np.uint8)
    oracle().ora_nv12_to_bgr(_ptr(out), _ptr(nv12), iw, ih, ws, nv12.shape[-1])
    return out


def rects_from_ptr(p, free=True):
    """copy a malloc()ed rect_t list (element 0 = header) into a numpy structured array of the real entries"""
    L = oracle()
    n = C.cast(p, C.POINTER(C.c_int))[0]
    buf = (C.c_char * (176 * n)).from_address(p)
    arr = np.frombuffer(bytes(buf), dtype=RECT_DTYPE).copy()
    if free:
        L.ora_free(p)
    return arr[1:]


class OracleRect:
    """the oracle's oclrect_t (oclrect.c:41-135): persistent buffers, executeOnce / gpu_task / cpu_task"""

    def __init__(self, iw, ih):
        self.L = oracle()
        self.iw, self.ih = iw, ih
        self.h = self.L.ora_rect_create(iw, ih)

    def close(self):
        if self.h:
            self.L.ora_rect_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def gpu_task(self, img, ws=None, stop_step=0):
        img = np.ascontiguousarray(img)
        self.L.ora_rect_gpu_task(self.h, _ptr(img), ws or img.shape[-1], stop_step)

    def cpu_task(self, tan_aov):
        return rects_from_ptr(self.L.ora_rect_cpu_task(self.h, tan_aov))

    def execute_once(self, img, tan_aov, ws=None):
        self.gpu_task(img, ws)
        return self.cpu_task(tan_aov)

    def buffer(self, name, dtype=np.int32):
        p = self.L.ora_rect_buffer(self.h, name.encode())
        n = self.iw * self.ih * (4 if name.startswith("ioBig") else 1)
        raw = (C.c_int32 * n).from_address(p)
        return np.frombuffer(raw, dtype=np.int32).view(dtype)

    def ls_list(self):
        raw = self.buffer("ioBig0")
        n = int(raw[0])
        return raw.view(np.uint8)[: 56 * (n + 1)].view(LS_DTYPE).copy()

    def times(self):
        t = np.zeros(5)
        self.L.ora_rect_last_times(self.h, _ptr(t))
        return t


def stats():
    s = OraStats()
    oracle().ora_get_stats(C.byref(s))
    return {n: getattr(s, n) for n, _ in OraStats._fields_}
