"""ctypes binding of ``include/whenet_b200.h``.  No CPU fallback: if the shared
library is missing or CUDA is unavailable every call raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))

PRECISIONS = {"fp32": 0, "bf16": 1, "fp16": 2}

EXPORTS = [
    "whenet_create", "whenet_load_weights", "whenet_export_packed", "whenet_import_packed", "whenet_set_stream", "whenet_forward_u8", "whenet_forward_u8_async", "whenet_forward_f32",
    "whenet_crop_resize_u8", "whenet_synchronize", "whenet_host_alloc", "whenet_host_free", "whenet_debug_enable_taps", "whenet_debug_tap",
    "whenet_debug_conv1x1", "whenet_debug_decode", "whenet_debug_raise_timeout", "whenet_debug_read_trace", "whenet_debug_set_k1_plan", "whenet_debug_set_k1w_plan", "whenet_profile_enable", "whenet_profile_read", "whenet_launch_count", "whenet_set_option",
    "whenet_last_error", "whenet_version", "whenet_destroy",
]


class WhenetError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("whenet_b200 error %d: %s" % (code, msg))
        self.code = code


class Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.POINTER(C.c_float)), ("ndim", C.c_int32), ("dims", C.c_int64 * 4)]


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("ms", C.c_float), ("launches", C.c_int), ("bytes", C.c_double),
                ("flops", C.c_double)]


def lib_path() -> str:
    return os.environ.get("WHENET_B200_LIB", os.path.join(HERE, "libwhenet_b200.so"))


_lib = None


def load():
    """dlopen the C-ABI library (building it first if the source is newer and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if "WHENET_B200_LIB" not in os.environ:
        from . import build
        try:
            build.build_lib()
        except FileNotFoundError:
            # no nvcc on this machine: a prebuilt library is acceptable.  A COMPILE or LINK failure is not - loading the
            # stale binary next to edited sources would run old kernels against new host code.
            if not os.path.exists(path):
                raise
    if not os.path.exists(path):
        raise WhenetError(-2, "shared library %s not found; run `python -m whenet_b200.build`" % path)
    L = C.CDLL(path)
    P = C.c_void_p
    L.whenet_create.argtypes = [C.POINTER(P), C.c_int, C.c_int, C.c_int]
    L.whenet_load_weights.argtypes = [P, C.POINTER(Tensor), C.c_int]
    L.whenet_export_packed.argtypes = [P, P, P, P, C.POINTER(C.c_int64)]
    L.whenet_import_packed.argtypes = [P, P, C.c_int64, P, C.c_int64, P, C.c_int64]
    L.whenet_set_stream.argtypes = [P, P]
    L.whenet_forward_u8.argtypes = [P, P, C.c_int, C.c_int, P, P, C.c_int]
    L.whenet_forward_f32.argtypes = [P, P, C.c_int, C.c_int, P, P, C.c_int]
    L.whenet_forward_u8_async.argtypes = [P, P, C.c_int, P, P]
    L.whenet_crop_resize_u8.argtypes = [P, P, C.c_int, C.c_int, C.c_int, P, C.c_int, C.c_int, P]
    L.whenet_synchronize.argtypes = [P]
    L.whenet_host_alloc.argtypes = [C.c_size_t]
    L.whenet_host_alloc.restype = P
    L.whenet_host_free.argtypes = [P]
    L.whenet_host_free.restype = None
    L.whenet_debug_enable_taps.argtypes = [P, C.c_int]
    L.whenet_debug_tap.argtypes = [P, C.c_char_p, P, C.c_size_t, C.POINTER(C.c_size_t)]
    L.whenet_debug_conv1x1.argtypes = [P, C.c_int, P, P, P, P, P, P, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int]
    L.whenet_debug_decode.argtypes = [P, P, C.c_int, P]
    L.whenet_debug_raise_timeout.argtypes = [P]
    L.whenet_debug_read_trace.argtypes = [P, P, C.c_int]
    L.whenet_debug_set_k1_plan.argtypes = [P] + [C.c_int] * 7
    L.whenet_debug_set_k1w_plan.argtypes = [P] + [C.c_int] * 8
    L.whenet_profile_enable.argtypes = [P, C.c_int]
    L.whenet_profile_read.argtypes = [P, C.POINTER(KernelStat), C.c_int]
    L.whenet_launch_count.argtypes = [P]
    L.whenet_launch_count.restype = C.c_int64
    L.whenet_set_option.argtypes = [P, C.c_char_p, C.c_int]
    L.whenet_last_error.restype = C.c_char_p
    L.whenet_version.restype = C.c_char_p
    L.whenet_destroy.argtypes = [P]
    L.whenet_destroy.restype = None
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise WhenetError(rc, load().whenet_last_error().decode("utf-8", "replace"))
