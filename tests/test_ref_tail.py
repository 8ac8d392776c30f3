"""The host tail (executeCPUTask) pinned to the REFERENCE's own code.

oracle/_ref/librd_ref_tail.so is the reference's oclrect.c + helper.c compiled here from /root/reference (oracle/Makefile
target _ref, oracle/ref_tail_wrap.c).  On the same inputs - the oracle's device-stage outputs - its executeCPUTask, the
oracle's restatement (ora_tail.cpp) and the product's tail (rd_tail.cpp, through the C-ABI entry rd_rect_tail, a host
function that needs no GPU) must produce the same rectangle list: same order, every field bit-exact.
tests/golden/ref_tail_golden.json holds reference-generated lists (tools/make_ref_tail_golden.py) for where the reference
sources and the prebuilt library are absent."""
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
import parity
import ref_tail_lib as rt
from tools_path import ROOT

FIELDS = ("c2", "c3", "value", "status")
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_tail_golden.json")))


def _device_outputs(iw, ih, seed):
    o = ol.OracleRect(iw, ih)
    o.gpu_task(ol.synth_frame(iw, ih, seed))
    arrays = tuple(np.ascontiguousarray(o.buffer(n)).copy() for n in ("ioBig0", "ioBig1", "iobuf1"))
    return o, arrays


def _product_tail(ls, votes, segid, iw, ih):
    import rectdetect_b200 as rd
    p = rd.lib().rd_rect_tail(ls.ctypes.data, segid.ctypes.data, votes.ctypes.data, iw, ih, parity.TAN_AOV)
    return rd.api.rects_from_ptr(p)


def _same(a, b):
    return len(a) == len(b) and all(np.ascontiguousarray(a[f]).tobytes() == np.ascontiguousarray(b[f]).tobytes() for f in FIELDS)


def _golden_rects(g):
    r = np.zeros(g["n_rects"], ol.RECT_DTYPE)
    r["status"] = g["status"]
    r["c2"] = np.array([float.fromhex(v) for v in g["c2"]]).reshape(-1, 4, 2)
    r["c3"] = np.array([float.fromhex(v) for v in g["c3"]]).reshape(-1, 4, 3)
    r["value"] = [float.fromhex(v) for v in g["value"]]
    return r


@pytest.mark.skipif(not rt.available(), reason="neither /root/reference nor a prebuilt oracle/_ref/librd_ref_tail.so")
@pytest.mark.parametrize("iw,ih,seed", [(640, 480, 1), (640, 480, 3), (333, 217, 7), (640, 360, 1002), (1280, 720, 1000)])
def test_reference_tail_equals_oracle_and_product_tails(iw, ih, seed):
    o, (ls, votes, segid) = _device_outputs(iw, ih, seed)
    ref = rt.execute_cpu_task(ls, votes, segid, iw, ih, parity.TAN_AOV)
    ora = o.cpu_task(parity.TAN_AOV)
    prod = _product_tail(ls, votes, segid, iw, ih)
    o.close()
    assert len(ref) > 0
    assert _same(ref, ora), "oracle tail differs from the reference's executeCPUTask"
    assert _same(ref, prod), "product tail differs from the reference's executeCPUTask"


@pytest.mark.parametrize("g", GOLDEN, ids=lambda g: "%dx%d-s%d" % (g["iw"], g["ih"], g["seed"]))
def test_tails_reproduce_reference_generated_fixture(g):
    # no reference code involved at run time: the committed lists came out of the reference's executeCPUTask
    o, (ls, votes, segid) = _device_outputs(g["iw"], g["ih"], g["seed"])
    want = _golden_rects(g)
    assert _same(want, o.cpu_task(parity.TAN_AOV))
    assert _same(want, _product_tail(ls, votes, segid, g["iw"], g["ih"]))
    o.close()
