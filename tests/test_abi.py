"""The C-ABI library loads and exports every symbol include/*.h declares; host-only entry points behave (CPU only,
no compute call that needs a device)."""
import ctypes as C
import math
import os
import re
import subprocess

import numpy as np

import oracle_lib as ol
from tools_path import ROOT


def _declared_symbols():
    names = set()
    for hdr in ("rectdetect_b200.h", os.path.join("CL", "cl.h")):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}()]*(?:\([^()]*\)[^;{}()]*)*\)\s*;", src):
            n = m.group(1)
            if n not in ("defined", "sizeof"):
                names.add(n)
    return names


def test_library_exports_every_declared_symbol(rd):
    L = rd.lib()
    declared = _declared_symbols()
    assert len(declared) > 90, sorted(declared)
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, "declared in include/ but not exported: %s" % missing
    for must in ("init_oclrect", "oclrect_executeOnce", "oclrect_enqueueTask", "oclrect_pollTask", "oclpolyline_execute",
                 "oclimgutil_iirblur_f_f", "oclimgutil_label8x_int_int", "clCreateBuffer", "simpleGetDevice", "loadPlan"):
        assert must in declared


# every function the reference's five public headers declare (oclhelper.h, helper.h, oclimgutil.h, oclpolyline.h, oclrect.h), pinned
# here so that the check also runs where /root/reference is absent; when it is present the list itself is checked against the headers
REFERENCE_API = """clStrError checkError ce getDeviceName simpleGetDevice simpleGetDevices simpleCreateContext simpleBuildProgram simpleSetKernelArg
runKernel1D runKernel2D runKernel1Dx runKernel2Dx waitForEvent clearPlan loadPlan savePlan startProfiling finishProfiling showPlan allocatePinnedMemory
freePinnedMemory getNextKernelID exitf readFileAsStr readFileAsStrN currentTimeMillis sleepMillis String_trim initArrayMap ArrayMap_dispose ArrayMap_size
ArrayMap_remove ArrayMap_put ArrayMap_get ArrayMap_keyArray ArrayMap_valueArray ArrayMap_getKey ArrayMap_getValue init_oclimgutil dispose_oclimgutil
oclimgutil_clear oclimgutil_copy oclimgutil_cast_i_f oclimgutil_cast_c_i oclimgutil_threshold_i_i oclimgutil_threshold_f_f oclimgutil_rand
oclimgutil_convert_bgr_luminancef oclimgutil_convert_bgr_lumaf oclimgutil_convert_bgr_labeli oclimgutil_edge_f_f oclimgutil_edgevec_f2_f
oclimgutil_thinthres_f_f_f2 oclimgutil_thincubic_f_f_f2 oclimgutil_label8x_int_int oclimgutil_iirblur_f_f oclimgutil_convert_plab_bgr
oclimgutil_convert_bgr_plab oclimgutil_unpack_f_f_f_plab oclimgutil_pack_plab_f_f_f oclimgutil_edgevec_f2_plab oclimgutil_edge_f_plab
oclimgutil_calcStrength oclimgutil_filterStrength init_oclpolyline dispose_oclpolyline oclpolyline_execute init_oclrect dispose_oclrect
oclrect_executeOnce oclrect_enqueueTask oclrect_pollTask""".split()


def test_library_exports_the_reference_headers_completely(rd):
    L = rd.lib()
    missing = [n for n in REFERENCE_API if not hasattr(L, n)]
    assert not missing, "declared by the reference's headers but not exported: %s" % missing
    ref = "/root/reference"
    if os.path.exists(os.path.join(ref, "oclrect.h")):
        found = set()
        for hdr in ("oclhelper.h", "helper.h", "oclimgutil.h", "oclpolyline.h", "oclrect.h"):
            src = open(os.path.join(ref, hdr)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            src = re.sub(r"//[^\n]*", "", src)
            for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}()]*\)\s*;", src):
                found.add(m.group(1))
        assert found == set(REFERENCE_API), (sorted(found - set(REFERENCE_API)), sorted(set(REFERENCE_API) - found))


def test_helper_h_text_utilities(rd, tmp_path):                                # helper.c:40-101
    L = rd.lib()
    L.readFileAsStr.restype, L.readFileAsStr.argtypes = C.c_void_p, [C.c_char_p, C.c_int]
    L.readFileAsStrN.restype, L.readFileAsStrN.argtypes = C.c_void_p, [C.POINTER(C.c_char_p)]
    L.String_trim.argtypes = [C.c_char_p]
    a, b = tmp_path / "a.txt", tmp_path / "b.txt"
    a.write_text("first\nfile\n")
    b.write_text("second")
    p = L.readFileAsStr(str(a).encode(), 1000)
    assert C.string_at(p) == b"first\nfile\n"
    L.rd_free(p)
    names = (C.c_char_p * 3)(str(a).encode(), str(b).encode(), None)
    p = L.readFileAsStrN(names)
    assert C.string_at(p) == b"first\nfile\nsecond"
    L.rd_free(p)
    buf = C.create_string_buffer(b"  \t padded text \n ")
    L.String_trim(buf)
    assert buf.value == b"padded text"
    L.ArrayMap_getKey.restype, L.ArrayMap_getKey.argtypes = C.c_uint64, [C.c_void_p, C.c_int]
    L.initArrayMap.restype = C.c_void_p
    L.ArrayMap_put.restype, L.ArrayMap_put.argtypes = C.c_void_p, [C.c_void_p, C.c_uint64, C.c_void_p]
    L.ArrayMap_keyArray.restype, L.ArrayMap_keyArray.argtypes = C.POINTER(C.c_uint64), [C.c_void_p]
    m = L.initArrayMap()
    for k in (5, 1029, 77, 2053):
        L.ArrayMap_put(m, k, 1000 + k)
    keys = L.ArrayMap_keyArray(m)
    assert [L.ArrayMap_getKey(m, i) for i in range(4)] == [keys[i] for i in range(4)]


def test_library_is_built_for_sm_100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "rectdetect_b200", "librectdetect_b200.so")], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_layouts_match_the_reference_wire_formats(rd):
    assert rd.LS_DTYPE.itemsize == 56 and rd.RECT_DTYPE.itemsize == 176      # oclpolyline.h:74-83, oclrect.h:5-15
    assert rd.RECT_DTYPE.fields["c3"][1] == 64 and rd.RECT_DTYPE.fields["value"][1] == 160 and rd.RECT_DTYPE.fields["status"][1] == 168


def test_arraymap_semantics(rd):                                              # helper.c:124-267
    L = rd.lib()
    for n, (res, args) in {"initArrayMap": (C.c_void_p, []), "ArrayMap_put": (C.c_void_p, [C.c_void_p, C.c_uint64, C.c_void_p]),
                           "ArrayMap_get": (C.c_void_p, [C.c_void_p, C.c_uint64]), "ArrayMap_remove": (C.c_void_p, [C.c_void_p, C.c_uint64]),
                           "ArrayMap_size": (C.c_int, [C.c_void_p]), "ArrayMap_keyArray": (C.POINTER(C.c_uint64), [C.c_void_p]),
                           "ArrayMap_dispose": (None, [C.c_void_p])}.items():
        getattr(L, n).restype, getattr(L, n).argtypes = res, args
    m = L.initArrayMap()
    keys = [5, 1029, 3, 1 << 40, 2053]            # 5, 1029 and 2053 do not share a bucket: the hash folds the high bits in
    for k in keys:
        assert L.ArrayMap_put(m, k, k + 100) is None
    assert L.ArrayMap_put(m, 3, 999) == 103 and L.ArrayMap_get(m, 3) == 999
    assert L.ArrayMap_size(m) == 5 and L.ArrayMap_get(m, 77) is None
    ka = L.ArrayMap_keyArray(m)
    got = [ka[i] for i in range(5)]
    bucket = lambda k: (k ^ (k >> 10) ^ (k >> 20) ^ (k >> 30)) & 1023
    assert got == sorted(keys, key=lambda k: (bucket(k), keys.index(k)))      # bucket order, then insertion order
    assert L.ArrayMap_remove(m, 5) == 105 and L.ArrayMap_size(m) == 4
    assert L.ArrayMap_put(m, 1029, None) == 1129 and L.ArrayMap_size(m) == 3   # put(NULL) removes
    L.ArrayMap_dispose(m)


def test_no_device_is_reported_not_emulated(rd):
    # in the build container there is no GPU: the library must say so (0 devices), not fall back to a CPU path
    assert rd.device_count() >= 0
    assert rd.lib().loadPlan(b"plan.txt", None) == 0          # rect.cpp:86 skips its autotune sweep
    assert rd.lib().currentTimeMillis() > 1_600_000_000_000


def test_host_tail_matches_oracle_bit_exact(rd):
    # executeCPUTask (oclrect.c:1049): the product's host tail against the oracle's, on the oracle's device-stage outputs
    tan_aov = math.tan(math.radians(36.0))
    for iw, ih, seed in ((640, 480, 1), (1280, 720, 2), (333, 217, 7)):
        img = ol.synth_frame(iw, ih, seed)
        o = ol.OracleRect(iw, ih)
        want = o.execute_once(img, tan_aov)
        got = rd.rect_tail(o.ls_list(), o.buffer("iobuf1"), o.buffer("ioBig1"), iw, ih, tan_aov)
        assert len(want) > 0 and want.tobytes() == got.tobytes()
        o.close()


def test_host_tail_empty_and_dead_lists(rd):
    iw, ih = 64, 48
    ls = np.zeros(1, rd.LS_DTYPE)                                             # header only, n = 0
    seg = np.full(iw * ih, -1, np.int32)
    votes = np.zeros(iw * ih * 4, np.int32)
    assert len(rd.rect_tail(ls, seg, votes, iw, ih, 1.0)) == 0
    ls = np.zeros(4, rd.LS_DTYPE)
    ls.view(np.int32)[0] = 3                                                  # three dead entries (polyid == 0)
    assert len(rd.rect_tail(ls, seg, votes, iw, ih, 1.0)) == 0
