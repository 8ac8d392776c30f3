"""ctypes binding of oracle/_ref/librd_ref_tail.so: the REFERENCE's own executeCPUTask (oclrect.c:1049-1226) compiled from
/root/reference by `make -C oracle _ref` (see oracle/ref_tail_wrap.c).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle_lib import RECT_DTYPE, ROOT

REF_SO = os.path.join(ROOT, "oracle", "_ref", "librd_ref_tail.so")
REF_SRC = "/root/reference/oclrect.c"
_lib = None


def available():
    """the library exists already (it travels with the repository snapshot) or can be built here"""
    return os.path.exists(REF_SO) or os.path.exists(REF_SRC)


def lib():
    global _lib
    if _lib is None:
        if os.path.exists(REF_SRC):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_ref"], stdout=subprocess.DEVNULL)
        L = C.CDLL(REF_SO)
        L.rd_ref_execute_cpu_task.restype = C.c_void_p
        L.rd_ref_execute_cpu_task.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_int, C.c_double]
        L.rd_ref_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def execute_cpu_task(ls, votes, segid, iw, ih, tan_aov):
    """ls: int32 view of the segment list (ioBig0), votes: vote table (ioBig1), segid: region map (iobuf1) -> rect array"""
    L = lib()
    ls, votes, segid = (np.ascontiguousarray(a, np.int32) for a in (ls, votes, segid))
    p = L.rd_ref_execute_cpu_task(ls.ctypes.data, votes.ctypes.data, segid.ctypes.data, iw, ih, tan_aov)
    n = C.cast(p, C.POINTER(C.c_int))[0]
    arr = np.frombuffer(C.string_at(p, 176 * n), dtype=RECT_DTYPE).copy()
    L.rd_ref_free(p)
    arr = arr[1:]
    arr["_pad"] = 0                    # the reference leaves the struct padding uninitialised
    return arr
