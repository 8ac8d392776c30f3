"""ctypes binding of oracle/_ref/librd_ref.so: THE REFERENCE ITSELF running on the host cores.

`make -C oracle _ref` compiles the reference's own host code (helper.c, oclhelper.c, oclimgutil.c, oclpolyline.c,
oclrect.c - unmodified, from /root/reference) together with its three OpenCL C kernel files compiled as C++
(oracle/cl_translate.py, oracle/cl_compat.h) and a synchronous host runtime (oracle/ref_cl_rt.cpp).  The entry points
bound here are the reference's own (init_oclimgutil, init_oclpolyline, init_oclrect, oclrect_executeOnce,
oclimgutil_*, oclpolyline_execute; oclrect.h:17-23, oclimgutil.h:74-100, oclpolyline.h:88) plus the few accessors of
oracle/ref_full_wrap.c.  Test infrastructure: tests/, tools/ and bench.py's reference / cpu_baseline legs only.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle_lib import LS_DTYPE, RECT_DTYPE, ROOT

REF_SO = os.path.join(ROOT, "oracle", "_ref", "librd_ref.so")
REF_SRC = "/root/reference/oclrect.c"
_lib = None
_default_ctx = None
vp, ci, cf, cd = C.c_void_p, C.c_int, C.c_float, C.c_double


def available():
    """the library exists already (it travels with the repository snapshot) or can be built here"""
    return os.path.exists(REF_SO) or os.path.exists(REF_SRC)


def build():
    """(re)build oracle/_ref from /root/reference where that exists (this container); elsewhere the prebuilt files are used.
    RD_REF_NO_BUILD=1 skips it (set for the worker processes of bench.py, whose parent has built already)."""
    if os.path.exists(REF_SRC) and not os.environ.get("RD_REF_NO_BUILD"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_ref"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(REF_SO)
        sig = {
            "simpleGetDevice": (vp, [ci]), "simpleCreateContext": (vp, [vp]), "clCreateCommandQueue": (vp, [vp, vp, C.c_uint64, vp]),
            "clCreateBuffer": (vp, [vp, C.c_uint64, C.c_size_t, vp, vp]), "clReleaseMemObject": (ci, [vp]),
            "init_oclimgutil": (vp, [vp, vp]), "init_oclpolyline": (vp, [vp, vp]), "init_oclrect": (vp, [vp, vp, vp, vp, vp, ci, ci]),
            "dispose_oclrect": (None, [vp]), "dispose_oclimgutil": (None, [vp]), "dispose_oclpolyline": (None, [vp]),
            "oclrect_executeOnce": (vp, [vp, vp, ci, cd]), "oclrect_enqueueTask": (None, [vp, vp, ci]), "oclrect_pollTask": (vp, [vp, cd]),
            "rd_ref_rect_buffer": (vp, [vp, C.c_char_p]), "rd_ref_rect_reset": (None, [vp]), "rd_ref_gen_gpu_task": (None, [vp, vp, ci]), "rd_ref_cpu_task": (vp, [vp, cd]),
            "rd_ref_mem_ptr": (vp, [vp]), "rd_ref_free": (None, [vp]),
            "rd_ref_set_threads": (None, [ci]), "rd_ref_get_threads": (ci, []), "rd_ref_launches": (C.c_long, []),
            "rd_ref_trace_reset": (None, []), "rd_ref_trace_name": (C.c_char_p, [C.c_long]), "rd_ref_set_launch_limit": (None, [C.c_long]),
            # L2 operators (oclimgutil.h:74-100): (thiz, cl_mem..., scalars..., queue, events) -> cl_event (NULL when events == NULL)
            "oclimgutil_clear": (vp, [vp, vp, ci, vp, vp]),
            "oclimgutil_cast_i_f": (vp, [vp, vp, vp, cf, ci, vp, vp]), "oclimgutil_cast_c_i": (vp, [vp, vp, vp, ci, vp, vp]),
            "oclimgutil_threshold_i_i": (vp, [vp, vp, vp, ci, ci, ci, ci, vp, vp]),
            "oclimgutil_threshold_f_f": (vp, [vp, vp, vp, cf, cf, cf, ci, vp, vp]),
            "oclimgutil_convert_plab_bgr": (vp, [vp, vp, vp, ci, ci, ci, vp, vp]),
            "oclimgutil_unpack_f_f_f_plab": (vp, [vp, vp, vp, vp, vp, ci, ci, vp, vp]),
            "oclimgutil_pack_plab_f_f_f": (vp, [vp, vp, vp, vp, vp, ci, ci, vp, vp]),
            "oclimgutil_iirblur_f_f": (vp, [vp, vp, vp, vp, vp, ci, ci, ci, vp, vp]),
            "oclimgutil_edgevec_f2_f": (vp, [vp, vp, vp, ci, ci, vp, vp]), "oclimgutil_edge_f_plab": (vp, [vp, vp, vp, ci, ci, vp, vp]),
            "oclimgutil_thinthres_f_f_f2": (vp, [vp, vp, vp, vp, ci, ci, vp, vp]), "oclimgutil_thincubic_f_f_f2": (vp, [vp, vp, vp, vp, ci, ci, vp, vp]),
            "oclimgutil_edgevec_f2_plab": (vp, [vp, vp, vp, ci, ci, vp, vp]), "oclimgutil_edge_f_f": (vp, [vp, vp, vp, ci, ci, vp, vp]),
            "oclimgutil_convert_bgr_plab": (vp, [vp, vp, vp, ci, ci, ci, vp, vp]), "oclimgutil_convert_bgr_lumaf": (vp, [vp, vp, vp, cf, ci, ci, ci, vp, vp]),
            "oclimgutil_convert_bgr_labeli": (vp, [vp, vp, vp, ci, ci, ci, ci, vp, vp]),
            "oclimgutil_label8x_int_int": (vp, [vp, vp, vp, vp, ci, ci, ci, vp, vp]),
            "oclimgutil_calcStrength": (vp, [vp, vp, vp, vp, ci, ci, vp, vp]),
            "oclimgutil_filterStrength": (vp, [vp, vp, vp, ci, ci, ci, vp, vp]),
            "oclpolyline_execute": (vp, [vp, vp, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, cf, ci, ci, ci, vp, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def set_threads(n):
    """1 (default): every NDRange runs its work-items in raster order on the calling thread; n > 1: rows over n threads"""
    lib().rd_ref_set_threads(n)


def kernel_direct(tag, name):
    """refcl_<tag>_<name>(gw, gh, kernel args...): one of the reference's kernels over an NDRange, on caller arrays"""
    return getattr(lib(), "refcl_%s_%s" % (tag, name))


def _rects(p):
    n = C.cast(p, C.POINTER(C.c_int))[0]
    arr = np.frombuffer(C.string_at(p, 176 * n), dtype=RECT_DTYPE).copy()
    lib().rd_ref_free(p)
    arr = arr[1:]
    arr["_pad"] = 0                    # the reference leaves the struct padding uninitialised
    return arr


class RefContext:
    """device / context / queue / oclimgutil / oclpolyline exactly as rect.cpp:60-70 sets them up"""

    def __init__(self):
        L = lib()
        self.L = L
        self.device = L.simpleGetDevice(0)
        self.context = L.simpleCreateContext(self.device)
        self.queue = L.clCreateCommandQueue(self.context, self.device, 0, None)
        self.imgutil = L.init_oclimgutil(self.device, self.context)
        self.polyline = L.init_oclpolyline(self.device, self.context)
        self._mems = []

    def mem(self, nbytes, init=None):
        """clCreateBuffer(CL_MEM_READ_WRITE [| CL_MEM_COPY_HOST_PTR]) -> cl_mem (poly.cpp:92-103)"""
        if init is not None:
            init = np.ascontiguousarray(init)
            assert init.nbytes <= nbytes
            m = self.L.clCreateBuffer(self.context, 1, nbytes, None, None)
            C.memmove(self.L.rd_ref_mem_ptr(m), init.ctypes.data, init.nbytes)
        else:
            m = self.L.clCreateBuffer(self.context, 1, nbytes, None, None)
        self._mems.append(m)
        return m

    def view(self, m, count, dtype=np.int32):
        """numpy view of the first `count` items of a cl_mem (host runtime: the buffer lives in host memory)"""
        dt = np.dtype(dtype)
        raw = (C.c_char * (count * dt.itemsize)).from_address(self.L.rd_ref_mem_ptr(m))
        return np.frombuffer(raw, dtype=dt)

    def release(self):
        for m in self._mems:
            self.L.clReleaseMemObject(m)
        self._mems = []


_pool = {}


class RefRect:
    """the reference's oclrect_t (oclrect.c:41-135) - init_oclrect / oclrect_executeOnce / enqueue / poll.
    The reference hands out at most 1000 kernel ids per process (oclhelper.c KERNELIDMAX; 19 per init_oclrect), so closed
    objects go back to a pool and come out again with every device plane zeroed (= the state after init_oclrect)."""

    def __init__(self, iw, ih, ctx=None):
        global _default_ctx
        if ctx is None:
            ctx = _default_ctx = _default_ctx or RefContext()
        self.ctx = ctx
        self.L = self.ctx.L
        self.iw, self.ih = iw, ih
        free = _pool.setdefault((id(self.ctx), iw, ih), [])
        if free:
            self.h = free.pop()
            self.L.rd_ref_rect_reset(self.h)
        else:
            self.h = self.L.init_oclrect(self.ctx.imgutil, self.ctx.polyline, self.ctx.device, self.ctx.context, self.ctx.queue, iw, ih)

    def close(self):
        if self.h:
            _pool[(id(self.ctx), self.iw, self.ih)].append(self.h)
            self.h = None

    def execute_once(self, img, tan_aov, ws=None):
        img = np.ascontiguousarray(img).copy()          # executeOnce copies the page back into the caller's image (oclrect.c:1241)
        return _rects(self.L.oclrect_executeOnce(self.h, img.ctypes.data, ws or img.shape[-1] * (img.shape[-2] if img.ndim == 3 else 1), tan_aov))

    def gpu_task(self, img, ws, launch_limit=-1):
        """genGPUTask alone; launch_limit = n stops the schedule after its n-th kernel launch (later launches are skipped)"""
        img = np.ascontiguousarray(img)
        self.L.rd_ref_trace_reset()
        self.L.rd_ref_set_launch_limit(launch_limit)
        self.L.rd_ref_gen_gpu_task(self.h, img.ctypes.data, ws)
        self.L.rd_ref_set_launch_limit(-1)

    def trace(self):
        return [self.L.rd_ref_trace_name(i).decode() for i in range(self.L.rd_ref_launches())]

    def cpu_task(self, tan_aov):
        return _rects(self.L.rd_ref_cpu_task(self.h, tan_aov))

    def buffer(self, name, dtype=np.int32):
        p = self.L.rd_ref_rect_buffer(self.h, name.encode())
        n = self.iw * self.ih * (4 if name.startswith("ioBig") else 1)
        raw = (C.c_int32 * n).from_address(p)
        return np.frombuffer(raw, dtype=np.int32).view(dtype)

    def ls_list(self):
        raw = self.buffer("ioBig0")
        n = int(raw[0])
        return raw.view(np.uint8)[: 56 * (n + 1)].view(LS_DTYPE).copy()
