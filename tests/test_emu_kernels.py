"""Host replays of CUDA kernels whose logic is shared with the host through a header (CPU tests, no GPU needed).

1. The exact despeckle2 kernels: kd2_pre / kd2_seq of rectdetect_b200/csrc/rd_despeckle2.cu are built from
the pure functions in rd_despeckle2.cuh (static record, merge of the row above, threshold map, composition of two maps);
tests/emu_despeckle2x.cpp replays the kernels' structure around them - per-row compaction, 32-entry chunks, the doubling scan over
the lanes, the carry between chunks, the two-row buffer.  What the replay produces must equal the oracle's raster-order
despeckle2 (= the reference kernel run sequentially, tests/test_ref_device.py) on pipeline planes of every frame size and on
adversarial random planes (long runs of small regions, ties, zero and negative sizes) - a check of the formulation that
needs no GPU (the -m gpu tests then check the kernels themselves).

2. The device tail (executeCPUTask on the GPU, rectdetect_b200/csrc/rd_gtail.cu): its logic is rd_gtail.cuh; tests/emu_gtail.cpp runs
the per-item kernels as loops and the warp-per-candidate kernel with its 32 lanes as fibers (every ballot / shuffle / syncwarp
is a rendezvous, so a missing synchronisation shows up as a wrong answer).  The rect_t lists must equal the oracle's tail - and
through it the reference's own executeCPUTask (tests/test_ref_tail.py) - byte for byte, in the same order.

3. The first pass of labelMergeMain as a row wavefront (rd_merge1.cuh, kernels k_m1_pre / k_m1_wave of rd_ccl.cu):
tests/emu_merge1.cpp runs every row as a lane, all rows in lock step M1_SKEW pixels apart, with the A / B split of the label
plane and the one-step operand prefetch of the kernel.  Result = the oracle's sequential pass (= the reference kernel,
tests/test_ref_device.py), and no address is touched by two lanes in one step when one of them writes."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from tools_path import ROOT

SO = os.path.join(ROOT, "tests", "_emu", "libemu_d2x.so")


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-o", SO, os.path.join(ROOT, "tests", "emu_despeckle2x.cpp")])
    E = C.CDLL(SO)
    E.emu_despeckle2x.argtypes = [C.c_void_p] * 3 + [C.c_int] * 3
    E.emu_despeckle2x.restype = C.c_long
    return E


def P(a):
    return a.ctypes.data


@pytest.mark.parametrize("iw,ih,seed", [(640, 480, 2), (641, 479, 32), (333, 217, 7), (130, 97, 31), (48, 40, 32), (96, 64, 31), (257, 511, 33), (1280, 720, 33)])
def test_replay_equals_the_oracle_on_pipeline_planes(emu, iw, ih, seed):
    LO = ol.oracle()
    n = iw * ih
    img = ol.synth_frame(iw, ih, seed)
    o = ol.OracleRect(iw, ih)
    o.gpu_task(img, img.shape[-1], 17)
    lab, size = o.buffer("buf5").copy(), o.buffer("tmp0").copy()
    o.close()
    LO.ora_rect_calcSize(P(size), P(lab), iw, ih)
    want = lab.copy()
    LO.ora_rect_despeckle2(P(want), P(size), 16, iw, ih)
    got = np.full(n, -99, np.int32)
    chunks = emu.emu_despeckle2x(P(got), P(lab), P(size), 16, iw, ih)
    assert np.array_equal(got, want)
    assert (want != lab).any() and chunks >= ih        # every row has at least its two frame pixels


@pytest.mark.parametrize("iw,ih,nlab,thre,seed", [(97, 61, 40, 16, 1), (64, 64, 6, 3, 2), (200, 9, 500, 1, 3), (33, 150, 12, 50, 4), (1, 40, 5, 2, 5),
                                                  (70, 1, 9, 2, 6), (128, 96, 3000, 2, 7)])
def test_replay_equals_the_oracle_on_random_planes(emu, iw, ih, nlab, thre, seed):
    """labels drawn from a small alphabet in blobs, sizes arbitrary (ties, zeros, negatives): long chains of small pixels in
    every direction, which is where the order of evaluation matters most"""
    LO = ol.oracle()
    rng = np.random.default_rng(seed)
    n = iw * ih
    ids = rng.choice(n, size=min(nlab, n), replace=False).astype(np.int32)
    coarse = rng.integers(0, len(ids), ((ih + 3) // 4, (iw + 3) // 4))
    lab = ids[np.kron(coarse, np.ones((4, 4), np.int64))[:ih, :iw]]
    noise = rng.random((ih, iw)) < 0.3
    lab = np.where(noise, ids[rng.integers(0, len(ids), (ih, iw))], lab).astype(np.int32).ravel()
    size = np.zeros(n, np.int32)
    size[ids] = rng.integers(-2, 3 * thre, len(ids))
    want = lab.copy()
    LO.ora_rect_despeckle2(P(want), P(size), thre, iw, ih)
    got = np.full(n, -99, np.int32)
    emu.emu_despeckle2x(P(got), P(lab), P(size), thre, iw, ih)
    assert np.array_equal(got, want)


# ------------------------------------------------------------------------------------------------ device tail
GT_SO = os.path.join(ROOT, "tests", "_emu", "libemu_gtail.so")


@pytest.fixture(scope="module")
def emu_tail():
    os.makedirs(os.path.dirname(GT_SO), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-o", GT_SO,
                           os.path.join(ROOT, "tests", "emu_gtail.cpp")])
    E = C.CDLL(GT_SO)
    E.emu_gtail.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p]
    return E


def _emu_tail(E, ls, seg, votes, iw, ih, tan):
    out, st = np.zeros(8192, ol.RECT_DTYPE), np.zeros(4, np.int32)
    n = E.emu_gtail(P(out), len(out), P(ls), P(seg), P(votes), iw, ih, tan, P(st))
    assert n >= 0, n
    return out[:n], st


@pytest.mark.parametrize("iw,ih,seed,dense,aov", [(640, 480, 1, 0, 36.0), (333, 217, 7, 0, 36.0), (641, 479, 32, 0, 25.0), (1280, 720, 1000, 0, 36.0),
                                                   (1280, 720, 7, 1, 36.0), (640, 360, 1001, 0, 50.0), (48, 40, 32, 0, 36.0)])
def test_device_tail_replay_equals_the_oracle_tail(emu_tail, iw, ih, seed, dense, aov):
    import math
    tan = math.tan(math.radians(aov))
    img = ol.dense_frame(iw, ih, seed, 5, 5) if dense else ol.synth_frame(iw, ih, seed)
    o = ol.OracleRect(iw, ih)
    want = o.execute_once(img, tan)
    ls, seg, votes = o.buffer("ioBig0").copy(), o.buffer("iobuf1").copy(), o.buffer("ioBig1").copy()
    o.close()
    got, st = _emu_tail(emu_tail, ls, seg, votes, iw, ih, tan)
    assert got.tobytes() == want.tobytes(), (len(got), len(want), st.tolist())
    if iw >= 333:
        assert len(want) > 0 and st[3] == len(want)


def test_device_tail_replay_with_vote_collisions_and_dead_segments(emu_tail):
    """a vote table whose slots are partly owned by other segments or empty (hash collisions, oclrect.c:1116-1121), dead list entries,
    and an empty list"""
    import math
    tan = math.tan(math.radians(36.0))
    iw, ih = 640, 480
    o = ol.OracleRect(iw, ih)
    o.execute_once(ol.dense_frame(iw, ih, 3, 3, 3), tan)
    ls, seg, votes = o.buffer("ioBig0").copy(), o.buffer("iobuf1").copy(), o.buffer("ioBig1").copy()
    o.close()
    rng = np.random.default_rng(11)
    v = votes.reshape(-1, 5)
    occ = np.flatnonzero(v[:, 0] > 0)
    v[occ[rng.random(len(occ)) < 0.15], 0] += 1                   # owned by "another segment": unclipped
    v[occ[rng.random(len(occ)) < 0.10]] = 0                       # empty slot: the pair contributes nothing
    n = int(ls[0])
    lsv = ls.view(np.uint8)[: 56 * (n + 1)].view(ol.LS_DTYPE)
    lsv["polyid"][1 + rng.choice(n, n // 10, replace=False)] = 0  # dead entries
    LO = ol.oracle()
    want = ol.rects_from_ptr(LO.ora_tail(P(ls), P(seg), P(votes), iw, ih, tan))
    got, st = _emu_tail(emu_tail, ls, seg, votes, iw, ih, tan)
    assert got.tobytes() == want.tobytes() and len(want) > 3
    ls[0] = 0
    got, _ = _emu_tail(emu_tail, ls, seg, votes, iw, ih, tan)
    assert len(got) == 0


# ---- 3. first pass of the merge labelling as a wavefront ----
SO_M1 = os.path.join(ROOT, "tests", "_emu", "libemu_m1.so")


@pytest.fixture(scope="module")
def emu_m1():
    os.makedirs(os.path.dirname(SO_M1), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-o", SO_M1, os.path.join(ROOT, "tests", "emu_merge1.cpp")])
    E = C.CDLL(SO_M1)
    E.emu_merge1.argtypes = [C.c_void_p] * 4 + [C.c_int] * 4
    E.emu_merge1.restype = C.c_long
    return E


def _m1_both(E, pix, mask, edge, iw, ih, check):
    want, got = np.zeros(iw * ih, np.int32), np.zeros(iw * ih, np.int32)
    ol.oracle().ora_rect_labelMerge_first_pass(P(want), P(pix), P(mask), P(edge), iw, ih)
    hazards = E.emu_merge1(P(got), P(pix), P(mask), P(edge), iw, ih, 4, check)
    return want, got, hazards


@pytest.mark.parametrize("iw,ih,seed", [(640, 480, 2), (641, 479, 33), (322, 200, 31), (130, 97, 31), (48, 40, 32), (1280, 720, 1000)])
def test_merge_first_pass_wavefront_on_pipeline_planes(emu_m1, iw, ih, seed):
    img = ol.synth_frame(iw, ih, seed)
    o = ol.OracleRect(iw, ih)
    o.gpu_task(img, img.shape[-1], 16)                               # up to mkMergeMask1: pix = buf4, mask = tmp1, edge = buf2
    pix, mask, edge = o.buffer("buf4").copy(), o.buffer("tmp1").copy(), o.buffer("buf2").copy()
    o.close()
    want, got, hazards = _m1_both(emu_m1, pix, mask, edge, iw, ih, 1 if iw <= 640 else 0)
    assert np.array_equal(want, got)
    assert hazards == 0


def test_merge_first_pass_wavefront_on_adversarial_planes(emu_m1):
    """few colours (long pointer chains, deep chases: the 8-step limit of the kernel's chase is exercised), dense masks (every
    pixel may adopt from every neighbour), dense and sparse edges, frames narrower than the skew and smaller than a warp"""
    rng = np.random.default_rng(7)
    for k in range(160):
        iw, ih = int(rng.integers(3, 100)), int(rng.integers(3, 80))
        if k % 5 == 0:
            iw, ih = int(rng.integers(3, 8)), int(rng.integers(30, 120))
        ncol = int(rng.integers(1, 4))
        pix = rng.integers(0, ncol + 1, iw * ih).astype(np.int32)
        if k % 3 == 0:                                               # horizontal / vertical stripes: chains as long as the frame
            yy, xx = np.divmod(np.arange(iw * ih), iw)
            pix = (((yy // int(rng.integers(1, 4))) + (xx // int(rng.integers(1, 30))) * (k % 2)) % (ncol + 1)).astype(np.int32)
        mask = (rng.random(iw * ih) < rng.choice([0.0, 0.02, 0.3, 0.9])).astype(np.int32)
        edge = (rng.random(iw * ih) < rng.choice([0.0, 0.05, 0.3])).astype(np.int32)
        want, got, hazards = _m1_both(emu_m1, pix, mask, edge, iw, ih, 1)
        assert np.array_equal(want, got), (k, iw, ih)
        assert hazards == 0, (k, iw, ih, hazards)
