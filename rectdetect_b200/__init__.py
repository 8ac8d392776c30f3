"""rectdetect_b200 - B200-native (sm_100a) implementation of the shibatch/rectdetect per-frame hot path.

The product is the C-ABI shared library librectdetect_b200.so (sources in csrc/); `api` binds it with ctypes.
"""
from . import api  # noqa: F401
from .api import LS_DTYPE, RECT_DTYPE, Batch, Device, Mem, OclRect, RectLists, device_count, kernel_launches, lib, rect_tail, rect_tail_device, set_merge_replay, get_merge_replay  # noqa: F401
