"""Host replay of the exact despeckle2 kernels (CPU): kd2_pre / kd2_seq of rectdetect_b200/csrc/rd_despeckle2.cu are built from
the pure functions in rd_despeckle2.cuh (static record, merge of the row above, threshold map, composition of two maps);
tests/emu_despeckle2x.cpp replays the kernels' structure around them - per-row compaction, 32-entry chunks, the doubling scan over
the lanes, the carry between chunks, the two-row buffer.  What the replay produces must equal the oracle's raster-order
despeckle2 (= the reference kernel run sequentially, tests/test_ref_device.py) on pipeline planes of every frame size and on
adversarial random planes (long runs of small regions, ties, zero and negative sizes) - a check of the formulation that
needs no GPU (the -m gpu tests then check the kernels themselves)."""
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
