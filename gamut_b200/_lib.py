"""ctypes loader for the C-ABI library (gamut_b200/libgamut_b200.so, declared in include/gamut_b200.h).

The library is the product; there is no CPU fallback. Importing works without a GPU (so that the
CPU test-suite can check the exported symbols), but every compute entry point fails loudly when no
sm_100 device is present.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, "libgamut_b200.so")

_lib = None


class GamutB200Error(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise GamutB200Error(
            f"{LIBPATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    L = C.CDLL(LIBPATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int, C.c_longlong, C.c_size_t
    L.gb200_init.restype = i32
    L.gb200_version.restype = C.c_char_p
    L.gb200_last_error.restype = C.c_char_p
    L.gb200_launch_count.restype = i64
    L.gb200_sm_count.restype = i32
    L.gb200_device_alloc.restype = vp
    L.gb200_device_alloc.argtypes = [sz]
    L.gb200_device_free.argtypes = [vp]
    L.gb200_host_alloc.restype = vp
    L.gb200_host_alloc.argtypes = [sz]
    L.gb200_host_free.argtypes = [vp]
    L.gb200_free.argtypes = [vp]
    L.gb200_copy_to_host.argtypes = [vp, vp, sz]
    L.gb200_copy_to_device.argtypes = [vp, vp, sz]
    L.gb200_pixel_type_size.argtypes = [i32]
    L.gb200_scanlines_inter_type.argtypes = [i32, i32]
    L.gb200_scanlines_convert.argtypes = [i32, vp, i32, i32, vp, i32, i32, i32]
    L.gb200_scanlines_convert_device.argtypes = [i32, vp, i64, i32, vp, i64, i32, i32, vp]
    _lib = L
    return L


def last_error() -> str:
    return lib().gb200_last_error().decode("utf-8", "replace")


def check(ok, what: str):
    if not ok:
        raise GamutB200Error(f"{what}: {last_error()}")
    return ok
