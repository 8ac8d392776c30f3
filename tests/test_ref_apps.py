"""The reference's demo programs, UNCHANGED, as the callers of the drop-in boundary: rect.cpp, poly.cpp,
vidrect.cpp (and vidpoly.cpp) are compiled from /root/reference by `make -C oracle _ref` against tests/opencv_stub (a stand-in
for the handful of OpenCV names they use) and include/CL/cl.h, and linked twice:

  oracle/_ref/apps/<app>_b200  against rectdetect_b200/librectdetect_b200.so - every symbol the programs need resolves, and on a
                               GPU box they RUN on the CUDA path (the -m gpu tests below);
  oracle/_ref/apps/<app>_ref   against oracle/_ref/librd_ref.so, the reference itself on the host (the CPU tests below).

The stand-in's line() logs every call, which is how the tests read what a program drew: rect / vidrect draw 4 sides + 2
diagonals per rectangle (cvPoint truncates the corner coordinates to int), poly draws every live polyline segment.
"""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
import parity
from tools_path import ROOT

APPS = os.path.join(ROOT, "oracle", "_ref", "apps")


def _have(app):
    return os.path.exists(os.path.join(APPS, app))


def write_ppm(path, img, iw, ih):
    rgb = np.ascontiguousarray(img.reshape(ih, -1)[:, : 3 * iw].reshape(ih, iw, 3)[:, :, ::-1])
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (iw, ih) + rgb.tobytes())


def write_stream(path, frames, iw, ih):
    with open(path, "wb") as f:
        f.write(b"RDV1 %d %d %d\n" % (iw, ih, len(frames)))
        for fr in frames:
            f.write(np.ascontiguousarray(fr.reshape(ih, -1)[:, : 3 * iw]).tobytes())


def run_app(app, args, cwd, timeout=600):
    log = os.path.join(cwd, "stub.log")
    if os.path.exists(log):
        os.remove(log)
    r = subprocess.run([os.path.join(APPS, app)] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True, timeout=timeout,
                       env=dict(os.environ, RD_STUB_LOG=log))
    assert r.returncode == 0, (app, r.returncode, r.stdout[-400:], r.stderr[-400:])
    frames, cur = [], []
    if os.path.exists(log):
        for ln in open(log):
            t = ln.split()
            if t[0] == "line":
                cur.append(tuple(int(v) for v in t[1:]))
            elif t[0] == "frame":
                frames.append(cur)
                cur = []
    return r.stdout, frames, cur


def rect_lines(rects, colours):
    """what showRect (rect.cpp:33-46) draws for a rect list: sides with the status colour / thickness, two 1-px diagonals"""
    out = []
    for r in rects:
        b, g, rr, th = colours[int(r["status"])]
        c = [(int(r["c2"][i][0]), int(r["c2"][i][1])) for i in range(4)]         # cvPoint(double, double): truncation
        for i in range(4):
            out.append(c[i] + c[(i + 1) % 4] + (b, g, rr, th))
        out.append(c[0] + c[2] + (b, g, rr, 1))
        out.append(c[1] + c[3] + (b, g, rr, 1))
    return out


RECT_COLOURS = {0: (255, 0, 0, 1), 2: (255, 0, 0, 1), 1: (0, 200, 255, 2), 3: (0, 0, 255, 2)}        # rect.cpp:112-124
VID_COLOURS = {0: (0, 255, 0, 1), 2: (255, 0, 0, 1), 1: (0, 200, 255, 2), 3: (0, 0, 255, 2)}         # vidrect.cpp:178-191


def same_drawing(got, want, slack=0):
    """two line lists as multisets; slack = allowed difference per coordinate (a corner at 217.9999 vs 218.0001 truncates apart)"""
    if len(got) != len(want):
        return False
    left = list(want)
    for g in got:
        hit = next((w for w in left if g[4:] == w[4:] and all(abs(a - b) <= slack for a, b in zip(g[:4], w[:4]))), None)
        if hit is None:
            return False
        left.remove(hit)
    return True


# ------------------------------------------------------------------------------------------ CPU: the programs on the reference
@pytest.mark.skipif(not _have("rect_ref"), reason="oracle/_ref/apps not built (needs /root/reference)")
def test_reference_rect_program_runs_on_the_reference_library(tmp_path):
    """rect.cpp end to end on librd_ref.so (its own autotune sweep rect.cpp:86-101 included, on a tiny frame first so that
    plan.txt exists): what it draws = the ORACLE's rectangles for the frame"""
    cwd = str(tmp_path)
    write_ppm(os.path.join(cwd, "small.ppm"), ol.synth_frame(96, 64, 1), 96, 64)
    out, _, _ = run_app("rect_ref", ["small.ppm", 0, "o.ppm"], cwd)
    assert "Creating plan" in out and os.path.exists(os.path.join(cwd, "plan.txt"))
    iw, ih, seed = 640, 480, 1
    img = ol.synth_frame(iw, ih, seed)
    write_ppm(os.path.join(cwd, "f.ppm"), img, iw, ih)
    out, _, lines = run_app("rect_ref", ["f.ppm", 0, "o.ppm"], cwd)
    assert "Creating plan" not in out
    o = ol.OracleRect(iw, ih)
    want = rect_lines(o.execute_once(img, parity.TAN_AOV), RECT_COLOURS)
    o.close()
    assert len(lines) >= 12 and same_drawing(lines, want, slack=1)
    assert os.path.getsize(os.path.join(cwd, "o.ppm")) > iw * ih * 3


@pytest.mark.skipif(not _have("poly_ref"), reason="oracle/_ref/apps not built (needs /root/reference)")
def test_reference_poly_program_runs_on_the_reference_library(tmp_path):
    """poly.cpp (config 1 of BASELINE.json: 640x480, edge -> polyline only) on librd_ref.so: the segments it draws are the
    oracle's polyline vertex list, coordinate for coordinate"""
    cwd = str(tmp_path)
    iw, ih, seed = 640, 480, 1
    img = ol.synth_frame(iw, ih, seed)
    write_ppm(os.path.join(cwd, "f.ppm"), img, iw, ih)
    _, _, lines = run_app("poly_ref", ["f.ppm", 0], cwd)
    assert same_drawing(lines, _poly_lines(img, iw, ih))
    assert os.path.exists(os.path.join(cwd, "output.png"))


@pytest.mark.skipif(not _have("vidpoly_ref"), reason="oracle/_ref/apps not built (needs /root/reference)")
def test_reference_vidpoly_program_runs_on_the_reference_library(tmp_path):
    """vidpoly.cpp (vidpoly.cpp:150-215: the poly pipeline per frame with strength 2000, minerror 1, sizeThre 10, all its
    device buffers reused from frame to frame without clearing) on librd_ref.so: every frame draws the segments the oracle
    computes for that frame from FRESH buffers - nothing in the L2 path depends on what an earlier frame left behind"""
    cwd = str(tmp_path)
    iw, ih, nf = 320, 240, 3
    frames = [ol.synth_frame(iw, ih, 4000 + i) for i in range(nf)]
    write_stream(os.path.join(cwd, "in.rdv"), frames, iw, ih)
    _, per_frame, _ = run_app("vidpoly_ref", [0, "in.rdv", "out.rdv"], cwd)
    assert len(per_frame) == nf
    for k in range(nf):
        L = _poly_list(frames[k], iw, ih, 1.0, 10, 2000)
        want = [(int(L["x0"][i]), int(L["y0"][i]), int(L["x1"][i]), int(L["y1"][i]), 255, 255, 255, 1) for i in range(1, len(L))]   # no polyid test (vidpoly.cpp:196-203)
        assert len(want) > 3 and same_drawing(per_frame[k], want), k


def _poly_list(img, iw, ih, minerror, size_thre, strength):
    n = iw * ih
    lsid, ls = np.zeros(n, np.int32), np.zeros(4 * n, np.int32)
    ol.oracle().ora_poly_frame(img.ctypes.data, img.shape[-1], iw, ih, minerror, size_thre, strength, lsid.ctypes.data, ls.ctypes.data, None)
    cnt = int(ls[0])
    return ls[: 14 * (cnt + 1)].view(np.uint8).view(ol.LS_DTYPE)


def _poly_lines(img, iw, ih):
    L = _poly_list(img, iw, ih, 1.0, 20, 500)
    cnt = len(L) - 1
    want = []
    for i in range(1, cnt + 1):                                                    # poly.cpp:138-154
        if L["polyid"][i] == 0 or L["leftPtr"][i] > 0:
            continue
        j, k = i, 0
        while j > 0:
            col = (100, 100, 255) if k & 1 else (255, 255, 100)
            want.append((int(L["x0"][j]), int(L["y0"][j]), int(L["x1"][j]), int(L["y1"][j])) + col + (1,))
            j, k = int(L["rightPtr"][j]), k + 1
    assert len(want) > 10
    return want


# ------------------------------------------------------------------------------------------ GPU: the programs on the CUDA library
@pytest.mark.gpu
@pytest.mark.skipif(not _have("rect_b200"), reason="oracle/_ref/apps not built (needs /root/reference at build time)")
def test_reference_rect_program_runs_unchanged_on_the_cuda_library(tmp_path):
    """rect.cpp, unmodified, linked against librectdetect_b200.so: config 2 of BASELINE.json (one 1280x720 frame, seed 2)"""
    cwd = str(tmp_path)
    iw, ih, seed = 1280, 720, 2
    img = ol.synth_frame(iw, ih, seed)
    write_ppm(os.path.join(cwd, "f.ppm"), img, iw, ih)
    out, _, lines = run_app("rect_b200", ["f.ppm", 0, "o.ppm"], cwd)
    assert "Creating plan" not in out                                             # loadPlan() == 0: no work-group autotuner on CUDA
    o = ol.OracleRect(iw, ih)
    want = rect_lines(o.execute_once(img, parity.TAN_AOV), RECT_COLOURS)
    o.close()
    assert len(lines) >= 12 and same_drawing(lines, want)


@pytest.mark.gpu
@pytest.mark.skipif(not _have("vidrect_b200"), reason="oracle/_ref/apps not built (needs /root/reference at build time)")
def test_reference_vidrect_program_runs_unchanged_on_the_cuda_library(tmp_path):
    """vidrect.cpp, unmodified (its enqueue / poll loop, vidrect.cpp:159-205), on a 6-frame 640x360 stream with AOV 72: per
    written frame the rectangles of the frame polled - the oracle object carries the strong-edge plane across frames (Q1)"""
    cwd = str(tmp_path)
    iw, ih, nf = 640, 360, 6
    frames = [ol.synth_frame(iw, ih, 3000 + i) for i in range(nf)]
    write_stream(os.path.join(cwd, "in.rdv"), frames, iw, ih)
    out, per_frame, _ = run_app("vidrect_b200", [0, "in.rdv", "out.rdv", 72], cwd)
    assert "Resolution : %d x %d" % (iw, ih) in out
    assert len(per_frame) == nf - 1                                               # the loop polls frame k after enqueueing k+1
    o = ol.OracleRect(iw, ih)
    for k in range(nf - 1):
        want = rect_lines(o.execute_once(frames[k], parity.TAN_AOV), VID_COLOURS)
        assert same_drawing(per_frame[k], want), k
    o.close()
    assert sum(len(f) for f in per_frame) > 0


@pytest.mark.gpu
@pytest.mark.skipif(not _have("poly_b200"), reason="oracle/_ref/apps not built (needs /root/reference at build time)")
def test_reference_poly_program_runs_unchanged_on_the_cuda_library(tmp_path):
    """poly.cpp, unmodified: L2 operators + raw clCreateBuffer / clEnqueueReadBuffer calls on the CUDA library (config 1)"""
    cwd = str(tmp_path)
    iw, ih, seed = 640, 480, 1
    img = ol.synth_frame(iw, ih, seed)
    write_ppm(os.path.join(cwd, "f.ppm"), img, iw, ih)
    _, _, lines = run_app("poly_b200", ["f.ppm", 0], cwd)
    assert same_drawing(lines, _poly_lines(img, iw, ih))
