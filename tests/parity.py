"""Step-by-step parity of the CUDA rect pipeline against the CPU oracle (test infrastructure).

compare_steps() runs the oracle's genGPUTask replay and the library's device schedule to the same stop step of
SURVEY.md section 10.1 on fresh objects and compares every buffer that step produces: integer planes bit-exact,
float planes bit-exact too (the canonical arithmetic is deterministic IEEE, so equality is expected - the 1e-4
tolerance of the spec is only needed for the final corner coordinates and is reported separately).
"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_lib as ol  # noqa: E402

# step -> [(buffer name, dtype, how many elements are meaningful: 'n', '2n', 'bytes_n', 'ls', '4n')]
STEP_BUFFERS = {
    1: [("buf0", np.uint32, "n")],
    # step 2 (unpack to three float planes) has no counterpart: the blur reads the packed plane directly
    3: [("tmp1", np.float32, "n"), ("tmp2", np.float32, "n"), ("tmp3", np.float32, "n")],
    4: [("buf1", np.uint32, "n")],
    5: [("ioBig0", np.float32, "2n")],
    6: [("tmp0", np.float32, "n")],
    7: [("buf1", np.float32, "n")],
    8: [("tmp1", np.int32, "n")],
    9: [("tmp1", np.int32, "n")],
    10: [("buf2", np.int32, "n")],
    11: [("buf2", np.int32, "n"), ("buf3", np.int32, "n")],
    12: [("tmp0", np.int32, "n"), ("tmp1", np.int8, "bytes_n")],
    13: [("buf4", np.uint32, "n")],
    14: [("buf4", np.uint32, "n")],
    15: [("buf2", np.int32, "n"), ("buf3", np.int32, "n")],
    16: [("tmp0", np.int32, "n"), ("tmp1", np.int32, "n")],
    17: [("buf5", np.int32, "n")],
    18: [("tmp0", np.int32, "n"), ("buf5", np.int32, "n")],
    19: [("tmp1", np.int32, "n"), ("iobuf1", np.int32, "n")],
    20: [("buf0", np.int32, "n"), ("tmp2", np.int32, "n"), ("ioBig0", np.int32, "ls")],
    21: [("ioBig1", np.int32, "4n")],
}


def _view(raw_i32, dtype, kind, n):
    b = raw_i32.view(np.uint8)
    if kind == "n":
        return b[: 4 * n].view(dtype)
    if kind == "2n":
        return b[: 8 * n].view(dtype)
    if kind == "4n":
        return b[: 16 * n].view(dtype)
    if kind == "bytes_n":
        return b[:n].view(dtype)
    if kind == "ls":
        cnt = int(raw_i32[0])
        return b[: 56 * (cnt + 1)].view(np.int32)
    raise ValueError(kind)


def _bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


def compare_steps(iw, ih, seed, steps, rd, dev, ws=None, frames_before=0):
    """returns list of (step, buffer, mismatches, total, detail)"""
    img = ol.synth_frame(iw, ih, seed, ws=ws)
    n = iw * ih
    out = []
    for k in steps:
        o = ol.OracleRect(iw, ih)
        g = rd.OclRect(dev, iw, ih)
        for f in range(frames_before):       # carry-over state: run earlier frames completely on both
            prev = ol.synth_frame(iw, ih, seed - frames_before + f, ws=ws)
            o.gpu_task(prev, ws=img.shape[-1])
            g.run_device(prev, stop_step=0)
        o.gpu_task(img, ws=img.shape[-1], stop_step=k)
        g.run_device(img, stop_step=k)
        for name, dtype, kind in STEP_BUFFERS[k]:
            a = _view(o.buffer(name), dtype, kind, n)
            b = _view(g.buffer(name), dtype, kind, n)
            if a.shape != b.shape:
                out.append((k, name, -1, a.size, "shape %s vs %s" % (a.shape, b.shape)))
                continue
            bad = np.flatnonzero(_bits(a) != _bits(b))
            detail = ""
            if bad.size:
                i = int(bad[0])
                detail = "first at %d (x=%d,y=%d): oracle %r cuda %r" % (i, i % iw, (i // iw) % ih, a[i], b[i])
            out.append((k, name, int(bad.size), int(a.size), detail))
        o.close()
        g.close()
    return out


# ---- the production schedule (rd_rect.cu : gpu_task_fast) stage by stage.  Its planes live in other buffers and narrower
# types than the reference's; every entry maps one plane to the oracle step / buffer that holds the same values.
# stage -> [(cuda buffer, dtype, kind, oracle step, oracle buffer, oracle dtype, oracle kind)]
FAST_STAGES = {
    1: [("aux0", np.uint32, "n", 1, "buf0", np.uint32, "n")],                         # packed Lab (the production schedule keeps it off the reference's plan)
    2: [("tmp1", np.float32, "n", 3, "tmp1", np.float32, "n"), ("tmp2", np.float32, "n", 3, "tmp2", np.float32, "n"),
        ("tmp3", np.float32, "n", 3, "tmp3", np.float32, "n"), ("buf1", np.uint32, "n", 4, "buf1", np.uint32, "n")],
    3: [("buf2", np.float32, "n", 7, "buf1", np.float32, "n")],                       # thinned edge strength
    4: [("tmp0", np.uint8, "bytes_n", 9, "tmp1", np.int32, "n")],                     # cleaned string image
    5: [("buf1", np.int32, "n", 10, "buf2", np.int32, "n")],                          # string components
    6: [("buf3", np.int32, "n", 11, "buf3", np.int32, "n"), ("tmp1", np.int8, "bytes_n", 12, "tmp1", np.int8, "bytes_n"),
        ("tmp5", np.int32, "n", 15, "buf3", np.int32, "n")],                          # strengths, weak mask, strong-edge bitmap
    7: [("buf4", np.uint32, "n", 13, "buf4", np.uint32, "n")],                        # edge-preserving blur
    8: [("tmp2", np.uint32, "n", 14, "buf4", np.uint32, "n")],                        # quantised + despeckled colours
    9: [("tmp3", np.uint8, "bytes_n", 16, "tmp1", np.int32, "n"), ("tmp0", np.int32, "n", 16, "tmp0", np.int32, "n")],
    10: [("buf4", np.int32, "n", 17, "buf5", np.int32, "n")],                         # colour regions
    11: [("tmp0", np.int32, "n", 18, "tmp0", np.int32, "n"), ("tmp1", np.int32, "n", 19, "tmp1", np.int32, "n")],
    12: [("iobuf1", np.int32, "n", 19, "iobuf1", np.int32, "n")],                     # segid map
    13: [("buf0", np.int32, "n", 20, "buf0", np.int32, "n"), ("ioBig0", np.int32, "ls", 20, "ioBig0", np.int32, "ls"),
         ("buf3", np.int32, "n", 15, "buf3", np.int32, "n")],
    0: [("ioBig1", np.int32, "4n", 21, "ioBig1", np.int32, "4n")],                    # vote table
}


def compare_fast_stages(iw, ih, seed, stages, rd, dev, ws=None):
    """production schedule stopped after each of `stages` (0 = run to the end) against the oracle's planes;
    returns list of (stage, buffer, mismatches, total, detail)"""
    img = ol.synth_frame(iw, ih, seed, ws=ws)
    n = iw * ih
    out = []
    ora = {}

    def oracle_plane(step, name, dtype, kind):
        if (step, name) not in ora:
            o = ol.OracleRect(iw, ih)
            o.gpu_task(img, ws=img.shape[-1], stop_step=step)
            for nm in ("buf0", "buf1", "buf2", "buf3", "buf4", "buf5", "tmp0", "tmp1", "tmp2", "tmp3", "iobuf1", "ioBig0", "ioBig1"):
                ora[(step, nm)] = o.buffer(nm).copy()
            o.close()
        return _view(ora[(step, name)], dtype, kind, n)

    for k in stages:
        g = rd.OclRect(dev, iw, ih)
        g.run_device(img, stop_step=-k)
        for name, dtype, kind, ostep, oname, odtype, okind in FAST_STAGES[k]:
            a = oracle_plane(ostep, oname, odtype, okind)
            b = _view(g.buffer(name), dtype, kind, n)
            if a.shape != b.shape:
                out.append((k, name, -1, a.size, "shape %s vs %s" % (a.shape, b.shape)))
                continue
            if a.dtype != b.dtype:
                a, b = a.astype(np.int64), b.astype(np.int64)
            bad = np.flatnonzero(_bits(a) != _bits(b))
            detail = ""
            if bad.size:
                i = int(bad[0])
                detail = "first at %d (x=%d,y=%d): oracle %r cuda %r" % (i, i % iw, (i // iw) % ih, a[i], b[i])
            out.append((k, name, int(bad.size), int(a.size), detail))
        g.close()
    return out


def canon_rects(r):
    """sort a rect list into a canonical order (the reference's order is a hash-map iteration order)"""
    if len(r) == 0:
        return r
    key = np.lexsort((r["c2"][:, 0, 1], r["c2"][:, 0, 0], r["status"]))
    return r[key]


def rects_close(a, b, rtol=1e-4):
    """set equality of two rect lists: status exact, corners within rtol (relative)"""
    if len(a) != len(b):
        return False, "count %d vs %d" % (len(a), len(b))
    a, b = canon_rects(a), canon_rects(b)
    if not np.array_equal(a["status"], b["status"]):
        return False, "status differs"
    for f in ("c2", "c3"):
        if not np.allclose(a[f], b[f], rtol=rtol, atol=1e-6):
            return False, "%s differs: max abs %g" % (f, float(np.abs(a[f] - b[f]).max()))
    return True, "ok (bit-exact: %s)" % (a.tobytes() == b.tobytes())


TAN_AOV = math.tan(math.radians(36.0))   # rect.cpp:84 : AOV 72 degrees
