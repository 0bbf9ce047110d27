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


class PngInfo(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("channels", C.c_int), ("file_channels", C.c_int),
                ("bits", C.c_int), ("ppmX", C.c_float), ("ppmY", C.c_float), ("pixelRatio", C.c_float)]


class QoiDesc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("channels", C.c_uint8), ("colorspace", C.c_uint8)]


class QoixDesc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("pitchBytes", C.c_int32), ("channels", C.c_uint8),
                ("bitdepth", C.c_uint8), ("colorspace", C.c_uint8), ("compression", C.c_uint8),
                ("pixelAspectRatio", C.c_float), ("resolutionY", C.c_float)]


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
        L.or_png_load.restype = C.c_void_p
        L.or_png_load.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(PngInfo)]
        L.or_png_is16.argtypes = [C.c_char_p, C.c_size_t]
        L.or_png_unfilter.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.or_jpeg_load.restype = C.c_void_p
        L.or_jpeg_load.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                   C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.or_qoi_decode.restype = C.c_void_p
        L.or_qoi_decode.argtypes = [C.c_char_p, C.c_int, C.POINTER(QoiDesc), C.c_int]
        L.or_qoix_lz4_decode.restype = C.c_void_p
        L.or_qoix_lz4_decode.argtypes = [C.c_char_p, C.c_int, C.POINTER(QoixDesc), C.c_int, C.POINTER(C.c_int)]
        L.or_qoix_lz4_encode.restype = C.c_void_p
        L.or_qoix_lz4_encode.argtypes = [C.c_void_p, C.POINTER(QoixDesc), C.c_int, C.POINTER(C.c_int)]
        L.or_qoi_encode.restype = C.c_void_p
        L.or_qoi_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.or_qoix_encode.restype = C.c_void_p
        L.or_qoix_encode.argtypes = [C.c_void_p, C.POINTER(QoixDesc), C.POINTER(C.c_int)]
        L.or_qoi10b_encode.restype = C.c_void_p
        L.or_qoi10b_encode.argtypes = [C.c_void_p, C.POINTER(QoixDesc), C.POINTER(C.c_int)]
        L.or_qoiplane_encode.restype = C.c_void_p
        L.or_qoiplane_encode.argtypes = [C.c_void_p, C.POINTER(QoixDesc), C.POINTER(C.c_int)]
        L.or_qoiplane10_encode.restype = C.c_void_p
        L.or_qoiplane10_encode.argtypes = [C.c_void_p, C.POINTER(QoixDesc), C.POINTER(C.c_int)]
        L.or_lz4_compress.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        L.or_lz4_compress_bound.argtypes = [C.c_int]
        L.or_lz4_decompress_fast.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        L.or_tga_load.restype = C.c_void_p
        L.or_tga_load.argtypes = [C.c_char_p, C.c_size_t] + [C.POINTER(C.c_int)] * 3
        L.or_tga_encode.restype = C.c_void_p
        L.or_tga_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.or_bmp_encode.restype = C.c_void_p
        L.or_bmp_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.POINTER(C.c_int)]
        L.or_bmp_load.restype = C.c_void_p
        L.or_bmp_load.argtypes = [C.c_char_p, C.c_size_t, C.c_int] + [C.POINTER(C.c_int)] * 3 + [C.POINTER(C.c_float)] * 3
        L.or_identify_format.argtypes = [C.c_char_p, C.c_size_t]
        L.or_zlib_decode.restype = C.c_void_p
        L.or_zlib_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_int, C.POINTER(C.c_size_t)]
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


def _take(ptr: int, nbytes: int) -> np.ndarray:
    """Copy a malloc'd oracle result into numpy and free it."""
    a = np.ctypeslib.as_array(C.cast(ptr, u8p), shape=(nbytes,)).copy() if nbytes else np.zeros(0, np.uint8)
    lib().or_free(ptr)
    return a


def png_load(data: bytes, req_comp: int = 0, want16: int = 0):
    """or_png_load. Returns (pixels (h, w, ch) uint8|uint16, info) or (None, info)."""
    L = lib()
    info = PngInfo()
    p = L.or_png_load(data, len(data), req_comp, want16, C.byref(info))
    if not p:
        return None, info
    n = info.width * info.height * info.channels * (2 if want16 else 1)
    a = _take(p, n)
    if want16:
        a = a.view(np.uint16)
    return a.reshape(info.height, info.width, info.channels), info


def png_is16(data: bytes) -> bool:
    return bool(lib().or_png_is16(data, len(data)))


def png_unfilter(raw: np.ndarray, img_n: int, out_n: int, w: int, h: int, depth: int):
    out = np.zeros(w * h * out_n * (2 if depth == 16 else 1), np.uint8)
    ok = lib().or_png_unfilter(raw.ctypes.data, raw.size, img_n, out_n, w, h, depth, out.ctypes.data)
    return out if ok else None


def zlib_decode(data: bytes, guess: int, parse_header: int = 1):
    n = C.c_size_t(0)
    p = lib().or_zlib_decode(data, len(data), guess, parse_header, C.byref(n))
    if not p:
        return None
    return _take(p, n.value)


def jpeg_load(data: bytes, req_comps: int = -1):
    """or_jpeg_load. Returns (pixels (h, w, c) uint8, actual_comps, par, dpiY) or None."""
    w, h, ac = C.c_int(), C.c_int(), C.c_int()
    par, dpi = C.c_float(), C.c_float()
    p = lib().or_jpeg_load(data, len(data), req_comps, C.byref(w), C.byref(h), C.byref(ac), C.byref(par), C.byref(dpi))
    if not p:
        return None
    c = ac.value if req_comps < 0 else req_comps
    a = _take(p, w.value * h.value * c)
    return a.reshape(h.value, w.value, c), ac.value, par.value, dpi.value


def qoi_decode(data: bytes, channels: int = 0):
    d = QoiDesc()
    p = lib().or_qoi_decode(data, len(data), C.byref(d), channels)
    if not p:
        return None
    c = channels if channels else d.channels
    return _take(p, d.width * d.height * c).reshape(d.height, d.width, c), d


def qoix_decode(data: bytes, flags: int = 0):
    """or_qoix_lz4_decode -> (pixels, desc, PixelType) or None."""
    d = QoixDesc()
    t = C.c_int(-1)
    p = lib().or_qoix_lz4_decode(data, len(data), C.byref(d), flags, C.byref(t))
    if not p:
        return None
    a = _take(p, d.pitchBytes * d.height)
    if d.bitdepth == 10:
        a = a.view(np.uint16)
    return a.reshape(d.height, d.width, d.channels), d, t.value


def qoix_encode(pixels: np.ndarray, bitdepth: int, colorspace: int = 0, force_lz4: bool = False,
                par: float = -1.0, dpi: float = -1.0) -> bytes:
    """or_qoix_lz4_encode of a (h, w, c) uint16 (10-bit streams) image."""
    h, w, c = pixels.shape
    px = np.ascontiguousarray(pixels)
    d = QoixDesc(w, h, w * c * px.itemsize, c, bitdepth, colorspace, 0, par, dpi)
    n = C.c_int(0)
    p = lib().or_qoix_lz4_encode(px.ctypes.data, C.byref(d), 1 if force_lz4 else 0, C.byref(n))
    if not p:
        raise RuntimeError("qoix encode failed")
    return _take(p, n.value).tobytes()


def qoi_encode(pixels: np.ndarray, colorspace: int = 0, pitch=None, first_scanline: int = 0, shape=None):
    """or_qoi_encode (qoi.d:295-426) of a (h, w, 3|4) uint8 image, or None. pitch / first_scanline / shape as in
    gamut_b200.codecs.qoi_encode (padded or vertically flipped storage)."""
    px = np.ascontiguousarray(pixels)
    h, w, c = shape if shape is not None else px.shape
    n = C.c_int(0)
    p = lib().or_qoi_encode(px.ctypes.data + first_scanline, w, h, pitch if pitch is not None else w * c, c, colorspace, C.byref(n))
    if not p:
        return None
    return _take(p, n.value).tobytes()


def qoi2avg_encode(pixels: np.ndarray, colorspace: int = 0, par: float = -1.0, dpi: float = -1.0, pitch=None, shape=None):
    """or_qoix_encode (qoi2avg.d:376-617) of a (h, w, 3|4) uint8 image: the QOI2AVG stream without the LZ4 stage, or None."""
    px = np.ascontiguousarray(pixels)
    h, w, c = shape if shape is not None else px.shape
    d = QoixDesc(w, h, pitch if pitch is not None else w * c, c, 8, colorspace, 0, par, dpi)
    n = C.c_int(0)
    p = lib().or_qoix_encode(px.ctypes.data, C.byref(d), C.byref(n))
    if not p:
        return None
    return _take(p, n.value).tobytes()


def qoi10b_encode(pixels: np.ndarray, colorspace: int = 0, par: float = -1.0, dpi: float = -1.0, pitch=None, shape=None):
    """or_qoi10b_encode (qoi10b.d:136-500) of a (h, w, 1..4) uint16 image: the QOI-10b stream (version 1) without the LZ4
    stage, or None."""
    px = np.ascontiguousarray(pixels)
    h, w, c = shape if shape is not None else px.shape
    d = QoixDesc(w, h, pitch if pitch is not None else w * c * 2, c, 10, colorspace, 0, par, dpi)
    n = C.c_int(0)
    p = lib().or_qoi10b_encode(px.ctypes.data, C.byref(d), C.byref(n))
    if not p:
        return None
    return _take(p, n.value).tobytes()


def qoiplane_encode(pixels: np.ndarray, colorspace: int = 0, par: float = -1.0, dpi: float = -1.0, pitch=None):
    """or_qoiplane_encode (qoiplane.d:109-375) of a (h, w, 1|2) uint8 image: the stream without the LZ4 stage, or None."""
    h, w, c = pixels.shape
    px = np.ascontiguousarray(pixels)
    d = QoixDesc(w, h, pitch if pitch is not None else w * c, c, 8, colorspace, 0, par, dpi)
    n = C.c_int(0)
    p = lib().or_qoiplane_encode(px.ctypes.data, C.byref(d), C.byref(n))
    if not p:
        return None
    return _take(p, n.value).tobytes()


def qoiplane10_encode(pixels: np.ndarray, colorspace: int = 0, par: float = -1.0, dpi: float = -1.0, pitch=None):
    """or_qoiplane10_encode (qoiplane10.d:99-314) of a (h, w, c) uint16 image: the stream without the LZ4 stage, or None."""
    h, w, c = pixels.shape
    px = np.ascontiguousarray(pixels)
    d = QoixDesc(w, h, pitch if pitch is not None else w * c * px.itemsize, c, 10, colorspace, 0, par, dpi)
    n = C.c_int(0)
    p = lib().or_qoiplane10_encode(px.ctypes.data, C.byref(d), C.byref(n))
    if not p:
        return None
    return _take(p, n.value).tobytes()


def lz4_compress(data: bytes) -> bytes:
    L = lib()
    out = np.zeros(L.or_lz4_compress_bound(len(data)) + 16, np.uint8)
    n = L.or_lz4_compress(data, out.ctypes.data, len(data))
    return out[:n].tobytes()


def lz4_decompress(data: bytes, orig: int):
    out = np.zeros(orig + 1, np.uint8)
    r = lib().or_lz4_decompress_fast(data + b"\0" * 16, out.ctypes.data, orig)
    return out[:orig] if r >= 0 else None


def bmp_load(data: bytes, req_comp: int = 0):
    """stbi_load_from_callbacks on a BMP (stbdec.d:2263): (pixels[h, w, c], file_comp, ppmX, ppmY, pixelRatio) or None."""
    x, y, comp = C.c_int(), C.c_int(), C.c_int()
    px, py, pr = C.c_float(), C.c_float(), C.c_float()
    p = lib().or_bmp_load(data, len(data), req_comp, C.byref(x), C.byref(y), C.byref(comp), C.byref(px), C.byref(py), C.byref(pr))
    if not p:
        return None
    c = req_comp if req_comp else comp.value
    n = x.value * y.value * c
    a = np.ctypeslib.as_array((C.c_uint8 * max(n, 1)).from_address(p))[:n].copy().reshape(y.value, x.value, c)
    lib().or_free(p)
    return a, comp.value, px.value, py.value, pr.value


def tga_load(data: bytes):
    """TGADecoder.getImageInfo + decodeImage (codecs/tga.d:313-588): pixels[h, w, c] (l8 / la8 / rgb8 / rgba8) or None."""
    x, y, comp = C.c_int(), C.c_int(), C.c_int()
    p = lib().or_tga_load(data, len(data), C.byref(x), C.byref(y), C.byref(comp))
    if not p:
        return None
    n = x.value * y.value * comp.value
    a = np.ctypeslib.as_array((C.c_uint8 * max(n, 1)).from_address(p))[:n].copy().reshape(y.value, x.value, comp.value)
    lib().or_free(p)
    return a


def tga_encode(pixels: np.ndarray, pitch=None, first_scanline: int = 0, shape=None, type_=None):
    """saveTGA -> TGAEncoder (plugins/tga.d:123-149, codecs/tga.d:62-292) of a (h, w, 1|2|3|4) uint8 image (l8 / la8 /
    rgb8 / rgba8): the run-length TGA file, or None."""
    px = np.ascontiguousarray(pixels)
    h, w, c = shape if shape is not None else px.shape
    t = type_ if type_ is not None else {1: 0, 2: 3, 3: 9, 4: 12}[c]
    n = C.c_int(0)
    p = lib().or_tga_encode(px.ctypes.data + first_scanline, t, w, h, pitch if pitch is not None else w * c, C.byref(n))
    if not p:
        return None
    return _take(p, n.value).tobytes()


def bmp_encode(pixels: np.ndarray, ppmX: float = -1.0, ppmY: float = -1.0, pitch=None, first_scanline: int = 0, shape=None, type_=None):
    """saveBMP -> write_bmp (plugins/bmp.d:166-194, codecs/bmpenc.d:25-113) of a (h, w, 3|4) uint8 image: the file (row
    padding zero), or None."""
    px = np.ascontiguousarray(pixels)
    h, w, c = shape if shape is not None else px.shape
    t = type_ if type_ is not None else {3: 9, 4: 12}.get(c, -1)
    n = C.c_int(0)
    p = lib().or_bmp_encode(px.ctypes.data + first_scanline, t, w, h, pitch if pitch is not None else w * c, ppmX, ppmY, C.byref(n))
    if not p:
        return None
    return _take(p, n.value).tobytes()


def identify_format(data: bytes) -> int:
    return int(lib().or_identify_format(data, len(data)))
