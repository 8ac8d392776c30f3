"""Host-side mirror of the reference's C interface for the rectangle-detection path, over librectdetect_b200.so.

The reference is a C library (oclimgutil.h, oclpolyline.h, oclrect.h, oclhelper.h); its callers are the C++ demo
programs rect.cpp / poly.cpp / vidrect.cpp.  This module binds the same entry points with ctypes - same names, same
argument order, same error behaviour (any failure inside the library prints to stderr and exits the process, as the
reference's exitf/ce do) - and adds thin Python conveniences (numpy <-> device buffers) for tests and bench.py.

There is no CPU fallback: loading fails loudly when the shared library is missing, and every entry point that
touches the device exits when no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librectdetect_b200.so")

CL_MEM_READ_WRITE = 1 << 0
CL_MEM_COPY_HOST_PTR = 1 << 5
CL_TRUE = 1

LS_DTYPE = np.dtype([("x0", "<f4"), ("y0", "<f4"), ("x1", "<f4"), ("y1", "<f4"),
                     ("startIndex", "<i4"), ("endIndex", "<i4"), ("leftPtr", "<i4"), ("rightPtr", "<i4"),
                     ("startCount", "<i4"), ("endCount", "<i4"), ("maxDist", "<i4"), ("polyid", "<i4"),
                     ("npix", "<i4"), ("level", "<i4")])                       # linesegment_t, oclpolyline.h:74-83
RECT_DTYPE = np.dtype([("c2", "<f8", (4, 2)), ("c3", "<f8", (4, 3)), ("value", "<f8"), ("status", "<u4"),
                       ("_pad", "<u4")])                                        # rect_t, oclrect.h:5-15
assert LS_DTYPE.itemsize == 56 and RECT_DTYPE.itemsize == 176

_lib = None


def lib():
    """the loaded C-ABI library (raises if it has not been built: python __graft_entry__.py build)"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("librectdetect_b200.so is missing (build it with `make -C rectdetect_b200/csrc`); "
                           "rectdetect_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i, f, d, sz, u64 = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_size_t, C.c_uint64
    op_tail = [vp, vp]      # cl_command_queue, const cl_event *
    sig = {
        # L1 (oclhelper.h / helper.h)
        "simpleGetDevice": (vp, [i]), "getDeviceName": (vp, [vp]), "simpleCreateContext": (vp, [vp]),
        "loadPlan": (i, [C.c_char_p, vp]), "waitForEvent": (None, [vp]),
        "allocatePinnedMemory": (vp, [sz, vp, vp]), "freePinnedMemory": (None, [vp, vp, vp]),
        "currentTimeMillis": (C.c_int64, []),
        # CL/cl.h subset
        "clCreateCommandQueue": (vp, [vp, vp, u64, vp]), "clReleaseCommandQueue": (i, [vp]), "clReleaseContext": (i, [vp]),
        "clFinish": (i, [vp]), "clCreateBuffer": (vp, [vp, u64, sz, vp, vp]), "clReleaseMemObject": (i, [vp]),
        "clReleaseEvent": (i, [vp]),
        "clEnqueueReadBuffer": (i, [vp, vp, C.c_uint, sz, sz, vp, C.c_uint, vp, vp]),
        "clEnqueueWriteBuffer": (i, [vp, vp, C.c_uint, sz, sz, vp, C.c_uint, vp, vp]),
        # L2 imgutil (oclimgutil.h:74-100)
        "init_oclimgutil": (vp, [vp, vp]), "dispose_oclimgutil": (None, [vp]),
        "oclimgutil_clear": (vp, [vp, vp, i] + op_tail), "oclimgutil_copy": (vp, [vp, vp, vp, i] + op_tail),
        "oclimgutil_cast_i_f": (vp, [vp, vp, vp, f, i] + op_tail), "oclimgutil_cast_c_i": (vp, [vp, vp, vp, i] + op_tail),
        "oclimgutil_threshold_i_i": (vp, [vp, vp, vp, i, i, i, i] + op_tail),
        "oclimgutil_threshold_f_f": (vp, [vp, vp, vp, f, f, f, i] + op_tail),
        "oclimgutil_rand": (vp, [vp, vp, i] + op_tail),
        "oclimgutil_convert_plab_bgr": (vp, [vp, vp, vp, i, i, i] + op_tail),
        "oclimgutil_unpack_f_f_f_plab": (vp, [vp, vp, vp, vp, vp, i, i] + op_tail),
        "oclimgutil_pack_plab_f_f_f": (vp, [vp, vp, vp, vp, vp, i, i] + op_tail),
        "oclimgutil_iirblur_f_f": (vp, [vp, vp, vp, vp, vp, i, i, i] + op_tail),
        "oclimgutil_edgevec_f2_f": (vp, [vp, vp, vp, i, i] + op_tail),
        "oclimgutil_edge_f_plab": (vp, [vp, vp, vp, i, i] + op_tail),
        "oclimgutil_thinthres_f_f_f2": (vp, [vp, vp, vp, vp, i, i] + op_tail), "oclimgutil_thincubic_f_f_f2": (vp, [vp, vp, vp, vp, i, i] + op_tail),
        "oclimgutil_edgevec_f2_plab": (vp, [vp, vp, vp, i, i] + op_tail), "oclimgutil_edge_f_f": (vp, [vp, vp, vp, i, i] + op_tail),
        "oclimgutil_convert_bgr_plab": (vp, [vp, vp, vp, i, i, i] + op_tail), "oclimgutil_convert_bgr_lumaf": (vp, [vp, vp, vp, f, i, i, i] + op_tail),
        "oclimgutil_convert_bgr_labeli": (vp, [vp, vp, vp, i, i, i, i] + op_tail),
        "oclimgutil_label8x_int_int": (vp, [vp, vp, vp, vp, i, i, i] + op_tail),
        "oclimgutil_calcStrength": (vp, [vp, vp, vp, vp, i, i] + op_tail),
        "oclimgutil_filterStrength": (vp, [vp, vp, vp, i, i, i] + op_tail),
        # L2 polyline (oclpolyline.h:85-88)
        "init_oclpolyline": (vp, [vp, vp]), "dispose_oclpolyline": (None, [vp]),
        "oclpolyline_execute": (vp, [vp, vp, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, f, i, i, i] + op_tail),
        # L3 (oclrect.h:17-23)
        "init_oclrect": (vp, [vp, vp, vp, vp, vp, i, i]), "dispose_oclrect": (None, [vp]),
        "oclrect_executeOnce": (vp, [vp, vp, i, d]), "oclrect_enqueueTask": (None, [vp, vp, i]), "oclrect_pollTask": (vp, [vp, d]),
        # extensions
        "rd_wrap_device_memory": (vp, [vp, sz]), "rd_mem_device_ptr": (vp, [vp]), "rd_mem_size": (sz, [vp]),
        "rd_wrap_stream": (vp, [vp, i]), "rd_queue_stream": (vp, [vp]), "rd_device_count": (i, []),
        "rd_version": (C.c_char_p, []), "rd_kernel_launches": (i, []), "rd_free": (None, [vp]),
        "rd_rect_simpleJunction": (None, [vp, vp, i, i, vp]), "rd_rect_simpleConnect": (None, [vp, vp, i, i, vp]),
        "rd_rect_stringify": (None, [vp, vp, i, i, i, vp]),
        "rd_rect_blblur0": (None, [vp, vp, vp, i, i, vp]), "rd_rect_blblur1": (None, [vp, vp, vp, i, i, vp]),
        "rd_rect_quantize": (None, [vp, vp, i, i, i, i, i, vp]), "rd_rect_despeckle": (None, [vp, vp, vp, i, i, vp]),
        "rd_rect_mkMergeMask0": (None, [vp, vp, i, i, vp]), "rd_rect_mkMergeMask1": (None, [vp, vp, i, i, vp]),
        "rd_set_merge_replay": (None, [i]),
        "rd_get_merge_replay": (i, []),
        "rd_rect_labelMerge": (None, [vp, vp, vp, vp, i, i, vp]),
        "rd_rect_calcSize": (None, [vp, vp, i, i, vp]), "rd_rect_despeckle2": (None, [vp, vp, vp, i, i, i, vp]),
        "rd_rect_markBoundary": (None, [vp, vp, i, i, vp]), "rd_rect_reduceLS": (None, [vp, vp, vp, i, i, i, vp]),
        "rd_oclrect_buffer": (vp, [vp, C.c_char_p]), "rd_oclrect_run_device": (None, [vp, vp, i, i]),
        "rd_batch_run_nv12": (None, [vp, vp, C.c_size_t, i, i, d, vp]), "rd_oclrect_executeOnceNV12": (vp, [vp, vp, i, d]),
        "rd_rect_lists_flatten": (vp, [vp, i, vp]),
        "rd_rect_tail": (vp, [vp, vp, vp, i, i, d]), "rd_rect_tail_device": (vp, [vp, vp, vp, i, i, d, vp]),
        "rd_batch_create": (vp, [i, i, i, i, i]), "rd_batch_destroy": (None, [vp]),
        "rd_batch_run": (None, [vp, vp, sz, i, i, d, vp]), "rd_batch_run_device": (None, [vp, vp, sz, i, i, d, vp]),
        "rd_batch_stage_ms": (None, [vp, vp]),
        "rd_profile_start": (None, [i, C.c_char_p]), "rd_profile_stop": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = None  # filled lazily by exported_symbols()


def device_count():
    return lib().rd_device_count()


def kernel_launches():
    return lib().rd_kernel_launches()


def set_merge_replay(on):
    """labelMergeMain: 0 = schedule-independent fixed point (default), 1 = replay the reference's first pass exactly, then the
    fixed point (include/rectdetect_b200.h); process-wide, set before creating OclRect / Batch objects"""
    lib().rd_set_merge_replay(1 if on else 0)


def get_merge_replay():
    return bool(lib().rd_get_merge_replay())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def profile_start(select=None, stages=False):
    """time every kernel (select=None) or only kernels whose name contains `select`, with CUDA events on their stream;
    stages=True prefixes every name with the stage of the production schedule ("A/kf_iir_h3")"""
    lib().rd_profile_start(3 if stages else (1 if select is None else 2), (select or "").encode())


def profile_stop():
    """-> {kernel name: (launches, total device ms)}"""
    out = {}
    for line in lib().rd_profile_stop().decode().splitlines():
        name, cnt, ms = line.rsplit(" ", 2)
        out[name] = (int(cnt), float(ms))
    return out


class RectLists:
    """n per-frame rect_t lists held as ONE flat array + offsets (a sequence of numpy views, made on demand): what a caller that
    handles thousands of lists per second wants instead of n small arrays"""

    def __init__(self, flat, counts):
        self.flat = flat
        self.offsets = np.concatenate([[0], np.cumsum(np.asarray(counts, np.int64))])

    def __len__(self):
        return len(self.offsets) - 1

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return self.flat[self.offsets[i]: self.offsets[i + 1]]

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def counts(self):
        return np.diff(self.offsets)


def rect_lists_from_ptrs(ptrs, n):
    """n malloc()ed rect_t lists (a ctypes array of pointers) -> RectLists (views of ONE flat copy); frees the lists"""
    L = lib()
    counts = np.zeros(n, np.int32)
    flat_p = L.rd_rect_lists_flatten(ptrs, n, _p(counts))
    total = int(counts.sum())
    flat = np.frombuffer(C.string_at(flat_p, 176 * total), dtype=RECT_DTYPE) if total else np.zeros(0, RECT_DTYPE)
    L.rd_free(flat_p)
    return RectLists(flat, counts)


def rects_from_ptr(p):
    """malloc()ed rect_t list (element 0 = header with nItems) -> numpy structured array of the real entries; frees p"""
    L = lib()
    n = C.cast(p, C.POINTER(C.c_int))[0]
    arr = np.frombuffer(C.string_at(p, 176 * n), dtype=RECT_DTYPE).copy()
    L.rd_free(p)
    return arr[1:]


class Device:
    """device + context + one in-order queue, created the way rect.cpp:60-64 does"""

    def __init__(self, did=0):
        L = lib()
        self.did = did
        self.device = L.simpleGetDevice(did)
        self.context = L.simpleCreateContext(self.device)
        self.queue = L.clCreateCommandQueue(self.context, self.device, 0, None)

    def finish(self):
        lib().clFinish(self.queue)

    def close(self):
        if self.queue:
            lib().clReleaseCommandQueue(self.queue)
            lib().clReleaseContext(self.context)
            self.queue = None

    # ---- buffers ----
    def buffer(self, nbytes=None, data=None):
        return Mem(self, nbytes=nbytes, data=data)


class Mem:
    """a cl_mem (device buffer) with numpy upload / download through clEnqueueWrite/ReadBuffer (poly.cpp:92-129)"""

    def __init__(self, dev, nbytes=None, data=None, handle=None):
        L = lib()
        self.dev = dev
        if handle is not None:
            self.h, self.nbytes, self.owned = handle, nbytes, False
            return
        if data is not None:
            data = np.ascontiguousarray(data)
            nbytes = max(nbytes or 0, data.nbytes)
        self.nbytes = nbytes
        zero = np.zeros(nbytes, np.uint8)
        if data is not None:
            zero[: data.nbytes] = data.reshape(-1).view(np.uint8)
        self.h = L.clCreateBuffer(dev.context, CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR, nbytes, _p(zero), None)
        self.owned = True

    def write(self, data):
        data = np.ascontiguousarray(data)
        lib().clEnqueueWriteBuffer(self.dev.queue, self.h, CL_TRUE, 0, data.nbytes, _p(data), 0, None, None)

    def read(self, dtype=np.int32, count=None, offset=0):
        dtype = np.dtype(dtype)
        n = (self.nbytes - offset) // dtype.itemsize if count is None else count
        out = np.empty(n, dtype)
        lib().clEnqueueReadBuffer(self.dev.queue, self.h, CL_TRUE, offset, out.nbytes, _p(out), 0, None, None)
        return out

    def release(self):
        if self.owned and self.h:
            lib().clReleaseMemObject(self.h)
            self.h = None


class OclRect:
    """oclrect_t (oclrect.h:17-23): init_oclrect / executeOnce / enqueueTask / pollTask / dispose"""

    def __init__(self, dev, iw, ih):
        L = lib()
        self.dev, self.iw, self.ih = dev, iw, ih
        self.imgutil = L.init_oclimgutil(dev.device, dev.context)
        self.polyline = L.init_oclpolyline(dev.device, dev.context)
        self.h = L.init_oclrect(self.imgutil, self.polyline, dev.device, dev.context, dev.queue, iw, ih)

    def execute_once(self, img, tan_aov, ws=None):
        img = np.ascontiguousarray(img)
        return rects_from_ptr(lib().oclrect_executeOnce(self.h, _p(img), ws or img.shape[-1], tan_aov))

    def execute_once_nv12(self, nv12, tan_aov, ystride=None):
        """nv12: (ih * 3 // 2, ystride) uint8 array - Y plane, then the interleaved UV plane"""
        nv12 = np.ascontiguousarray(nv12)
        return rects_from_ptr(lib().rd_oclrect_executeOnceNV12(self.h, _p(nv12), ystride or nv12.shape[-1], tan_aov))

    def enqueue_task(self, img, ws=None):
        img = np.ascontiguousarray(img)
        lib().oclrect_enqueueTask(self.h, _p(img), ws or img.shape[-1])

    def poll_task(self, tan_aov):
        return rects_from_ptr(lib().oclrect_pollTask(self.h, tan_aov))

    def run_device(self, img, ws=None, stop_step=0):
        img = np.ascontiguousarray(img)
        lib().rd_oclrect_run_device(self.h, _p(img), ws or img.shape[-1], stop_step)

    def buffer(self, name, dtype=np.int32):
        big = name.startswith("ioBig")
        m = Mem(self.dev, nbytes=self.iw * self.ih * (16 if big else 4), handle=lib().rd_oclrect_buffer(self.h, name.encode()))
        return m.read(dtype)

    def ls_list(self):
        raw = self.buffer("ioBig0", np.int32)
        n = int(raw[0])
        return raw.view(np.uint8)[: 56 * (n + 1)].view(LS_DTYPE).copy()

    def close(self):
        if self.h:
            L = lib()
            L.dispose_oclrect(self.h)
            L.dispose_oclpolyline(self.polyline)
            L.dispose_oclimgutil(self.imgutil)
            self.h = None


class Batch:
    """frame-batch engine: nctx pipeline objects (streams) on one device, each launch processes frames_per_launch frames;
    frames are independent"""

    def __init__(self, device, iw, ih, nctx=4, frames_per_launch=8):
        self.iw, self.ih = iw, ih
        self.h = lib().rd_batch_create(device, iw, ih, nctx, frames_per_launch)

    def run(self, frames_ptr, frame_stride, ws, nframes, tan_aov, on_device=False, want_rects=True):
        """frames_ptr: address of nframes BGR8 frames (host, ideally pinned, or device when on_device)"""
        L = lib()
        out = (C.c_void_p * nframes)() if want_rects else None
        if on_device:
            L.rd_batch_run_device(self.h, frames_ptr, frame_stride, ws, nframes, tan_aov, out)
        else:
            if out is None:
                out = (C.c_void_p * nframes)()
            L.rd_batch_run(self.h, frames_ptr, frame_stride, ws, nframes, tan_aov, out)
        if out is None:
            return None
        return rect_lists_from_ptrs(out, nframes)

    def run_nv12(self, frames_ptr, frame_stride, ystride, nframes, tan_aov):
        """NV12 frames in host or device memory (rd_batch_run_nv12)"""
        out = (C.c_void_p * nframes)()
        lib().rd_batch_run_nv12(self.h, frames_ptr, frame_stride, ystride, nframes, tan_aov, out)
        return rect_lists_from_ptrs(out, nframes)

    def stage_ms(self):
        """host-side accounting of the last run: (ms the driver threads waited for the device, ms of host-tail phases)"""
        v = (C.c_double * 5)()
        lib().rd_batch_stage_ms(self.h, v)
        return float(v[0]), float(v[4])

    def close(self):
        if self.h:
            lib().rd_batch_destroy(self.h)
            self.h = None


def rect_tail(ls, segid, votes, iw, ih, tan_aov):
    """executeCPUTask (oclrect.c:1049) on host arrays: pure host code, no device needed"""
    ls = np.ascontiguousarray(ls)
    segid = np.ascontiguousarray(segid, np.int32)
    votes = np.ascontiguousarray(votes, np.int32)
    return rects_from_ptr(lib().rd_rect_tail(_p(ls), _p(segid), _p(votes), iw, ih, tan_aov))


def rect_tail_device(dev, ls, segid, votes, iw, ih, tan_aov):
    """executeCPUTask (oclrect.c:1049) on the device (rd_gtail.cu) from host arrays: uploads them and runs the operator"""
    n = iw * ih
    mls, mseg, mvotes = dev.buffer(nbytes=16 * n), dev.buffer(nbytes=4 * n), dev.buffer(nbytes=16 * n)
    raw = np.zeros(4 * n, np.int32)
    src = np.ascontiguousarray(ls).view(np.int32).ravel()
    raw[: src.size] = src
    mls.write(raw)
    mseg.write(np.ascontiguousarray(segid, np.int32))
    mvotes.write(np.ascontiguousarray(votes, np.int32))
    r = rects_from_ptr(lib().rd_rect_tail_device(mls.h, mseg.h, mvotes.h, iw, ih, tan_aov, dev.queue))
    for m in (mls, mseg, mvotes):
        m.release()
    return r
