"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so). TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".c", ".h"))]
    stale = (not os.path.exists(LIBPATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIBPATH) for s in srcs)
    if force or stale:
        if not any(os.access(os.path.join(p, "gcc"), os.X_OK) for p in os.environ.get("PATH", "").split(os.pathsep)):
            if os.path.exists(LIBPATH):
                return LIBPATH
        r = subprocess.run(["make", "-C", HERE, "-B" if force else "-s"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return LIBPATH


_lib = None
u8p = C.POINTER(C.c_uint8)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.or_pixelTypeSize.argtypes = [C.c_int]
        L.or_scanlinesInterType.argtypes = [C.c_int, C.c_int]
        L.or_scanlinesConvert.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                          C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.or_scanlinesCopy.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.or_free.argtypes = [C.c_void_p]
    return _lib


def _ptr(a: np.ndarray, offset: int = 0) -> int:
    return a.ctypes.data + offset


def scanlines_convert(src_type: int, src: np.ndarray, src_pitch: int, dst_type: int, dst: np.ndarray,
                      dst_pitch: int, w: int, h: int, src_off: int = 0, dst_off: int = 0) -> bool:
    """or_scanlinesConvert on numpy byte buffers; *_off = byte offset of the first scanline."""
    L = lib()
    inter = L.or_scanlinesInterType(src_type, dst_type)
    interbuf = np.zeros(max(1, w) * 16, dtype=np.uint8)
    return bool(L.or_scanlinesConvert(src_type, _ptr(src, src_off), src_pitch, dst_type, _ptr(dst, dst_off),
                                      dst_pitch, w, h, inter, _ptr(interbuf)))
