"""Deterministic synthetic BGR frames (librd_synth.so, csrc/rd_synth.cpp): workload input for tests and bench.py."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SYNTH_SO = os.path.join(_HERE, "librd_synth.so")
_syn = None


def synth_lib():
    global _syn
    if _syn is None:
        if not os.path.exists(SYNTH_SO):
            subprocess.check_call(["/usr/bin/g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", SYNTH_SO, os.path.join(_HERE, "csrc", "rd_synth.cpp")])
        L = C.CDLL(SYNTH_SO)
        L.rd_synth_frame.restype = C.c_int
        L.rd_synth_frame.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_int]
        _syn = L
    return _syn


def synth_frame(iw, ih, seed, ws=None, with_truth=False, out=None):
    """BGR8 frame as a (ih, ws) uint8 array (ws defaults to 3*iw); a (seed, iw, ih) triple names a frame"""
    ws = ws or 3 * iw
    img = np.zeros((ih, ws), np.uint8) if out is None else out
    q = np.zeros((256, 8), np.float64)
    n = synth_lib().rd_synth_frame(img.ctypes.data_as(C.c_void_p), iw, ih, ws, seed, q.ctypes.data_as(C.c_void_p), 256)
    return (img, q[:n].copy()) if with_truth else img
