"""Python bindings of the codec-level C ABI (include/gamut_b200.h): the seam one level below the
reference's plugins (stbi_load_from_callbacks, decompress_jpeg_image_from_stream, qoi_decode,
qoix_lz4_decode). All pixel work happens in the CUDA library."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib

u8p = C.POINTER(C.c_uint8)


class ImageDesc(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_int), ("height", C.c_int), ("channels", C.c_int),
                ("file_channels", C.c_int), ("bits", C.c_int), ("pixel_type", C.c_int), ("pitch", C.c_int),
                ("status", C.c_int), ("ppmX", C.c_float), ("ppmY", C.c_float), ("pixelAspectRatio", C.c_float)]


class LoadedImage(C.Structure):     # gb200_image
    _fields_ = [("alloc", C.c_void_p), ("alloc_bytes", C.c_size_t), ("data", C.c_void_p), ("width", C.c_int),
                ("height", C.c_int), ("type", C.c_int), ("pitch", C.c_int), ("layout", C.c_int),
                ("pixelAspectRatio", C.c_float), ("resolutionY", C.c_float), ("error", C.c_char_p)]


class TgaDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("pitchBytes", C.c_int32), ("type", C.c_int32)]


class BmpDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("pitchBytes", C.c_int32), ("type", C.c_int32),
                ("ppmX", C.c_float), ("ppmY", C.c_float)]


class QoiDesc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("channels", C.c_uint8), ("colorspace", C.c_uint8)]


class QoixDesc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("pitchBytes", C.c_int32), ("channels", C.c_uint8),
                ("bitdepth", C.c_uint8), ("colorspace", C.c_uint8), ("compression", C.c_uint8),
                ("pixelAspectRatio", C.c_float), ("resolutionY", C.c_float)]


_declared = False


def _L():
    global _declared
    L = _lib.lib()
    if not _declared:
        vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
        fp = C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int)
        L.gb200_batch_count.argtypes = [vp]
        L.gb200_batch_images.restype = C.POINTER(ImageDesc)
        L.gb200_batch_images.argtypes = [vp]
        L.gb200_batch_free.argtypes = [vp]
        L.gb200_batch_download.argtypes = [vp, vp, sz]
        L.gb200_batch_timing.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_double)]
        L.gb200_png_is16.argtypes = [C.c_char_p, sz]
        L.gb200_png_load.restype = vp
        L.gb200_png_load.argtypes = [C.c_char_p, sz, i32, i32, ip, ip, ip, fp, fp, fp]
        L.gb200_png_decode_batch.restype = vp
        L.gb200_png_decode_batch.argtypes = [i32, C.POINTER(C.c_char_p), C.POINTER(sz), C.POINTER(vp), i32, i32, vp]
        L.gb200_png_unfilter_device.argtypes = [vp, sz, vp, sz, i32, i32, i32, i32, vp, vp]
        L.gb200_inflate_device.argtypes = [i32, C.POINTER(vp), C.POINTER(C.c_uint32), C.POINTER(vp),
                                           C.POINTER(C.c_uint32), i32, vp, vp, vp]
        L.gb200_inflate_set_mode.argtypes = [i32]
        L.gb200_inflate_set_mode.restype = None
        L.gb200_jpeg_load.restype = vp
        L.gb200_jpeg_load.argtypes = [C.c_char_p, sz, i32, ip, ip, ip, fp, fp]
        L.gb200_jpeg_decode_batch.restype = vp
        L.gb200_jpeg_decode_batch.argtypes = [i32, C.POINTER(C.c_char_p), C.POINTER(sz), C.POINTER(vp), i32, vp]
        L.gb200_qoi_decode.restype = vp
        L.gb200_qoi_decode.argtypes = [C.c_char_p, i32, C.POINTER(QoiDesc), i32]
        L.gb200_qoi_encode.restype = vp
        L.gb200_qoi_encode.argtypes = [vp, C.POINTER(QoiDesc), i32, ip]
        L.gb200_qoi_encode_bound.restype = sz
        L.gb200_qoi_encode_bound.argtypes = [C.POINTER(QoiDesc)]
        L.gb200_qoi_encode_batch_device.restype = i32
        L.gb200_qoi_encode_batch_device.argtypes = [i32, C.POINTER(vp), C.POINTER(QoiDesc), ip, C.POINTER(vp), ip, vp]
        L.gb200_qoix_decode.restype = vp
        L.gb200_qoix_decode.argtypes = [C.c_char_p, i32, C.POINTER(QoixDesc), i32, ip]
        L.gb200_qoix_encode.restype = vp
        L.gb200_qoix_encode.argtypes = [vp, C.POINTER(QoixDesc), ip]
        L.gb200_qoix_encode_bound.restype = sz
        L.gb200_qoix_encode_bound.argtypes = [C.POINTER(QoixDesc)]
        L.gb200_qoix_encode_batch_device.restype = i32
        L.gb200_qoix_encode_batch_device.argtypes = [i32, C.POINTER(vp), C.POINTER(QoixDesc), C.POINTER(vp), ip, vp]
        L.gb200_qoix_decode_batch.restype = vp
        L.gb200_qoix_decode_batch.argtypes = [i32, C.POINTER(C.c_char_p), C.POINTER(sz), C.POINTER(vp), i32, vp]
        L.gb200_jpeg_probe.argtypes = [C.c_char_p, sz]
        L.gb200_tga_load.restype = vp
        L.gb200_tga_load.argtypes = [C.c_char_p, sz, ip, ip, ip]
        L.gb200_tga_decode_batch.restype = vp
        L.gb200_tga_decode_batch.argtypes = [i32, C.POINTER(C.c_char_p), C.POINTER(sz), C.POINTER(vp), vp]
        L.gb200_tga_encode.restype = vp
        L.gb200_tga_encode.argtypes = [vp, C.POINTER(TgaDesc), ip]
        L.gb200_tga_encode_bound.restype = sz
        L.gb200_tga_encode_bound.argtypes = [C.POINTER(TgaDesc)]
        L.gb200_tga_encode_batch_device.restype = i32
        L.gb200_tga_encode_batch_device.argtypes = [i32, C.POINTER(vp), C.POINTER(TgaDesc), C.POINTER(vp), ip, vp]
        L.gb200_bmp_encode.restype = vp
        L.gb200_bmp_encode.argtypes = [vp, C.POINTER(BmpDesc), ip]
        L.gb200_bmp_encode_size.restype = sz
        L.gb200_bmp_encode_size.argtypes = [C.POINTER(BmpDesc)]
        L.gb200_bmp_load.restype = vp
        L.gb200_bmp_load.argtypes = [C.c_char_p, sz, i32, ip, ip, ip, fp, fp, fp]
        L.gb200_bmp_decode_batch.restype = vp
        L.gb200_bmp_decode_batch.argtypes = [i32, C.POINTER(C.c_char_p), C.POINTER(sz), C.POINTER(vp), i32, vp]
        L.gb200_identify_format.argtypes = [C.c_char_p, sz]
        L.gb200_image_load.argtypes = [C.c_char_p, sz, i32, C.POINTER(LoadedImage)]
        L.gb200_decode_batch_host.argtypes = [i32, i32, C.POINTER(C.c_char_p), C.POINTER(sz), i32, i32, vp, sz,
                                              C.POINTER(ImageDesc), i32]
        _declared = True
    return L


def _take_host(ptr: int, nbytes: int) -> np.ndarray:
    a = np.ctypeslib.as_array(C.cast(ptr, u8p), shape=(max(nbytes, 1),))[:nbytes].copy()
    _lib.lib().gb200_free(ptr)
    return a


@dataclass
class PngResult:
    pixels: np.ndarray      # (h, w, channels) uint8 or uint16
    width: int
    height: int
    file_channels: int
    ppmX: float
    ppmY: float
    pixelRatio: float


def png_is16(data: bytes) -> bool:
    return bool(_L().gb200_png_is16(data, len(data)))


def png_load(data: bytes, req_comp: int = 0, want16: bool = False) -> Optional[PngResult]:
    """stbi_load_from_callbacks / stbi_load_16_from_callbacks (stbdec.d:713-735) on a memory buffer."""
    L = _L()
    w, h, comp = C.c_int(), C.c_int(), C.c_int()
    px, py, pr = C.c_float(), C.c_float(), C.c_float()
    p = L.gb200_png_load(data, len(data), req_comp, 1 if want16 else 0, C.byref(w), C.byref(h), C.byref(comp),
                         C.byref(px), C.byref(py), C.byref(pr))
    if not p:
        return None
    ch = req_comp if req_comp else comp.value
    n = w.value * h.value * ch * (2 if want16 else 1)
    a = _take_host(p, n)
    if want16:
        a = a.view(np.uint16)
    return PngResult(a.reshape(h.value, w.value, ch), w.value, h.value, comp.value, px.value, py.value, pr.value)


class Batch:
    """Owner of a device-resident decoded batch (gb200_batch)."""

    def __init__(self, handle: int):
        self.handle = handle
        L = _L()
        n = L.gb200_batch_count(handle)
        arr = L.gb200_batch_images(handle)
        self.images = [arr[i] for i in range(n)]

    def to_host(self, i: int) -> Optional[np.ndarray]:
        d = self.images[i]
        if not d.status:
            return None
        n = d.pitch * d.height
        out = np.empty(n, np.uint8)
        _lib.check(_L().gb200_copy_to_host(out.ctypes.data, d.pixels, n), "copy_to_host")
        a = out.view(np.uint16) if d.bits == 16 else out
        return a.reshape(d.height, d.width, d.channels)

    def download(self, dst_host: int, stride: int) -> None:
        """All decoded images to (pinned) host memory, image i at dst_host + i*stride."""
        _lib.check(_L().gb200_batch_download(self.handle, dst_host, stride), "batch_download")

    def timing(self):
        ph = (C.c_float * 8)()
        hp = C.c_double()
        _L().gb200_batch_timing(self.handle, ph, C.byref(hp))
        return list(ph), hp.value

    def free(self):
        if self.handle:
            _L().gb200_batch_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _batch_args(files: Sequence[bytes], files_dev: Optional[Sequence[int]]):
    n = len(files)
    arr = (C.c_char_p * n)(*files)
    lens = (C.c_size_t * n)(*[len(f) for f in files])
    dev = (C.c_void_p * n)(*files_dev) if files_dev is not None else None
    return n, arr, lens, dev


def png_decode_batch(files: Sequence[bytes], req_comp: int = 0, want16: int = -1,
                     files_dev: Optional[Sequence[int]] = None, stream: int = 0) -> Batch:
    n, arr, lens, dev = _batch_args(files, files_dev)
    h = _L().gb200_png_decode_batch(n, arr, lens, dev, req_comp, want16, stream)
    if not h:
        raise _lib.GamutB200Error("png_decode_batch: " + _lib.last_error())
    return Batch(h)


def inflate_set_mode(parallel: bool) -> None:
    _L().gb200_inflate_set_mode(1 if parallel else 0)


def inflate_device(streams: Sequence[bytes], caps: Sequence[int], parse_header: bool = True):
    """Kernel-level inflate (replaces miniz mz_uncompress3 as called from stbdec.d:1267-1321) of device-resident
    streams. Returns [(status, out_len, bytes)] -- status 0 ok, 1 output buffer too small, 2 corrupt."""
    import torch
    L = _L()
    n = len(streams)
    ins, outs = [], []
    for s, cap in zip(streams, caps):
        t = torch.zeros(((len(s) + 3) // 4) * 4 + 32, dtype=torch.uint8, device="cuda")
        if len(s):
            t[:len(s)] = torch.frombuffer(bytearray(s), dtype=torch.uint8).cuda()
        ins.append(t)
        outs.append(torch.zeros(cap + 32, dtype=torch.uint8, device="cuda"))
    lens_d = torch.zeros(n, dtype=torch.int32, device="cuda")
    st_d = torch.zeros(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    inp = (C.c_void_p * n)(*[t.data_ptr() for t in ins])
    outp = (C.c_void_p * n)(*[t.data_ptr() for t in outs])
    il = (C.c_uint32 * n)(*[len(s) for s in streams])
    oc = (C.c_uint32 * n)(*caps)
    _lib.check(L.gb200_inflate_device(n, inp, il, outp, oc, 1 if parse_header else 0, lens_d.data_ptr(), st_d.data_ptr(), None),
               "inflate_device")
    lens = lens_d.cpu().numpy()
    sts = st_d.cpu().numpy()
    return [(int(sts[i]), int(lens[i]), outs[i][:int(lens[i])].cpu().numpy().tobytes()) for i in range(n)]


@dataclass
class JpegResult:
    pixels: np.ndarray      # (h, w, channels) uint8
    width: int
    height: int
    actual_comps: int
    pixelAspectRatio: float
    dotsPerInchY: float


def jpeg_load(data: bytes, req_comps: int = -1) -> Optional[JpegResult]:
    """decompress_jpeg_image_from_stream (jpegload.d:3720) on a memory buffer."""
    L = _L()
    w, h, ac = C.c_int(), C.c_int(), C.c_int()
    par, dpi = C.c_float(), C.c_float()
    p = L.gb200_jpeg_load(data, len(data), req_comps, C.byref(w), C.byref(h), C.byref(ac), C.byref(par), C.byref(dpi))
    if not p:
        return None
    c = ac.value if req_comps < 0 else req_comps
    a = _take_host(p, w.value * h.value * c)
    return JpegResult(a.reshape(h.value, w.value, c), w.value, h.value, ac.value, par.value, dpi.value)


def jpeg_probe(data: bytes) -> int:
    """0 decodable here, 1 progressive (SOF2), 2 non-interleaved multi-scan, -1 not a JPEG (gb200_jpeg_probe)."""
    return int(_L().gb200_jpeg_probe(data, len(data)))


def jpeg_decode_batch(files: Sequence[bytes], req_comps: int = -1, files_dev: Optional[Sequence[int]] = None,
                      stream: int = 0) -> Batch:
    n, arr, lens, dev = _batch_args(files, files_dev)
    h = _L().gb200_jpeg_decode_batch(n, arr, lens, dev, req_comps, stream)
    if not h:
        raise _lib.GamutB200Error("jpeg_decode_batch: " + _lib.last_error())
    return Batch(h)


def qoi_decode(data: bytes, channels: int = 0):
    """qoi_decode (qoi.d:448). Returns (pixels (h, w, c) uint8, desc) or None."""
    d = QoiDesc()
    p = _L().gb200_qoi_decode(data, len(data), C.byref(d), channels)
    if not p:
        return None
    c = channels if channels else d.channels
    return _take_host(p, d.width * d.height * c).reshape(d.height, d.width, c), d


def qoi_encode(pixels: np.ndarray, colorspace: int = 0, pitch: Optional[int] = None, first_scanline: int = 0,
               shape: Optional[tuple] = None) -> Optional[bytes]:
    """qoi_encode (qoi.d:295) as saveQOI calls it (plugins/qoi.d:150): a (h, w, 3|4) uint8 image -> the QOI file, or None
    where the reference returns null. `pitch` (bytes, may be negative), `first_scanline` (byte offset of the first
    scanline in `pixels`) and `shape` = (h, w, c) describe padded or vertically flipped storage."""
    px = np.ascontiguousarray(pixels)
    h, w, c = shape if shape is not None else px.shape
    d = QoiDesc(w, h, c, colorspace)
    n = C.c_int(0)
    p = _L().gb200_qoi_encode(px.ctypes.data + first_scanline, C.byref(d), pitch if pitch is not None else w * c, C.byref(n))
    if not p:
        return None
    return _take_host(p, n.value).tobytes()


def qoi_encode_bound(w: int, h: int, c: int) -> int:
    d = QoiDesc(w, h, c, 0)
    return int(_L().gb200_qoi_encode_bound(C.byref(d)))


def qoi_encode_batch_device(pixels_dev: Sequence[int], shapes: Sequence[tuple], out_dev: Sequence[int], stream: int = 0,
                            pitches: Optional[Sequence[int]] = None):
    """gb200_qoi_encode_batch_device: device pointers of (h, w, c) uint8 images (gapless unless `pitches` is given) ->
    device buffers; returns the stream lengths (0 = refused)."""
    n = len(pixels_dev)
    pin = (C.c_void_p * max(n, 1))(*pixels_dev)
    pout = (C.c_void_p * max(n, 1))(*out_dev)
    descs = (QoiDesc * max(n, 1))()
    pit = (C.c_int * max(n, 1))()
    for i, (h, w, c) in enumerate(shapes):
        descs[i] = QoiDesc(w, h, c, 0)
        pit[i] = pitches[i] if pitches is not None else w * c
    lens = (C.c_int * max(n, 1))()
    _lib.check(_L().gb200_qoi_encode_batch_device(n, pin, descs, pit, pout, lens, stream), "qoi_encode_batch_device")
    return [lens[i] for i in range(n)]


def qoix_decode(data: bytes, flags: int = 0):
    """qoix_lz4_decode (plugins/qoix.d:350). Returns (pixels (h, w, c) uint8|uint16, desc, PixelType) or None."""
    d = QoixDesc()
    t = C.c_int(-1)
    p = _L().gb200_qoix_decode(data, len(data), C.byref(d), flags, C.byref(t))
    if not p:
        return None
    a = _take_host(p, d.pitchBytes * d.height)
    if d.bitdepth == 10:
        a = a.view(np.uint16)
    return a.reshape(d.height, d.width, d.channels), d, t.value


def qoix_encode(pixels: np.ndarray, bitdepth: Optional[int] = None, colorspace: int = 0, par: float = -1.0, dpi: float = -1.0,
                pitch: Optional[int] = None) -> Optional[bytes]:
    """qoix_lz4_encode (plugins/qoix.d:251) of a (h, w, c) image: c = 1|2 uint16 (10-bit values expanded to 16 bits) -> the
    QOI-Plane10 stream (qoiplane10.d:99), c = 1|2 uint8 -> the QOI-Plane stream (qoiplane.d:109), c = 3|4 uint8 -> the
    QOI2AVG stream (qoi2avg.d:376), c = 3|4 uint16 -> the QOI-10b stream (qoi10b.d:136); never LZ4-wrapped. None if the
    encoder refuses the image."""
    h, w, c = pixels.shape
    px = np.ascontiguousarray(pixels)
    if bitdepth is None:
        bitdepth = 10 if px.itemsize == 2 else 8
    d = QoixDesc(w, h, pitch if pitch is not None else w * c * px.itemsize, c, bitdepth, colorspace, 0, par, dpi)
    n = C.c_int(0)
    p = _L().gb200_qoix_encode(px.ctypes.data, C.byref(d), C.byref(n))
    if not p:
        return None
    return _take_host(p, n.value).tobytes()


def qoix_encode_bound(w: int, h: int, c: int) -> int:
    d = QoixDesc(w, h, w * c * 2, c, 10, 0, 0, -1.0, -1.0)
    return int(_L().gb200_qoix_encode_bound(C.byref(d)))


def qoix_encode_batch_device(pixels_dev: Sequence[int], shapes: Sequence[tuple], out_dev: Sequence[int], stream: int = 0,
                             bitdepths: Optional[Sequence[int]] = None):
    """gb200_qoix_encode_batch_device: device pointers of gapless (h, w, c) images (uint16 for bitdepth 10, the default;
    uint8 for bitdepth 8) -> device buffers; returns the stream lengths."""
    n = len(pixels_dev)
    pin = (C.c_void_p * max(n, 1))(*pixels_dev)
    pout = (C.c_void_p * max(n, 1))(*out_dev)
    descs = (QoixDesc * max(n, 1))()
    for i, (h, w, c) in enumerate(shapes):
        bd = bitdepths[i] if bitdepths is not None else 10
        descs[i] = QoixDesc(w, h, w * c * (2 if bd == 10 else 1), c, bd, 0, 0, -1.0, -1.0)
    lens = (C.c_int * max(n, 1))()
    _lib.check(_L().gb200_qoix_encode_batch_device(n, pin, descs, pout, lens, stream), "qoix_encode_batch_device")
    return [lens[i] for i in range(n)]


def qoix_decode_batch(files: Sequence[bytes], flags: int = 0, files_dev: Optional[Sequence[int]] = None,
                      stream: int = 0) -> Batch:
    n, arr, lens, dev = _batch_args(files, files_dev)
    h = _L().gb200_qoix_decode_batch(n, arr, lens, dev, flags, stream)
    if not h:
        raise _lib.GamutB200Error("qoix_decode_batch: " + _lib.last_error())
    return Batch(h)


def decode_batch_host(fmt: int, files: Sequence[bytes], arg: int, want16: int, dst_host: int, dst_stride: int,
                      sub_batch: int = 0):
    """gb200_decode_batch_host: host files in, pixels in host memory at dst_host + i*dst_stride, transfers overlapped
    with the decode of the next sub-batch. Returns the list of descriptors (pixels = host address, 0 if failed)."""
    n, arr, lens, _ = _batch_args(files, None)
    descs = (ImageDesc * max(n, 1))()
    _lib.check(_L().gb200_decode_batch_host(int(fmt), n, arr, lens, arg, want16, dst_host, dst_stride, descs, sub_batch),
               "decode_batch_host")
    return [descs[i] for i in range(n)]


def tga_load(data: bytes):
    """TGADecoder.getImageInfo + decodeImage (codecs/tga.d:313-588), as plugins/tga.d:45-70 calls them: pixels
    (h, w, c) uint8 with c = 1 (l8), 2 (la8), 3 (rgb8) or 4 (rgba8), or None."""
    w, h, comp = C.c_int(0), C.c_int(0), C.c_int(0)
    p = _L().gb200_tga_load(data, len(data), C.byref(w), C.byref(h), C.byref(comp))
    if not p:
        return None
    return _take_host(p, w.value * h.value * comp.value).reshape(h.value, w.value, comp.value)


def tga_decode_batch(files: Sequence[bytes], files_dev: Optional[Sequence[int]] = None, stream: int = 0) -> Batch:
    n, arr, lens, dev = _batch_args(files, files_dev)
    h = _L().gb200_tga_decode_batch(n, arr, lens, dev, stream)
    if not h:
        raise _lib.GamutB200Error("tga_decode_batch: " + _lib.last_error())
    return Batch(h)


_TGA_TYPE = {1: 0, 2: 3, 3: 9, 4: 12}          # channels -> PixelType l8 / la8 / rgb8 / rgba8


def tga_encode(pixels: np.ndarray, pitch: Optional[int] = None, first_scanline: int = 0, shape: Optional[tuple] = None,
               type_: Optional[int] = None) -> Optional[bytes]:
    """saveTGA (plugins/tga.d:123-149) of a (h, w, 1|2|3|4) uint8 image (l8 / la8 / rgb8 / rgba8): the run-length TGA file
    TGAEncoder writes, or None where saveTGA fails. pitch / first_scanline / shape describe padded or flipped storage."""
    px = np.ascontiguousarray(pixels)
    h, w, c = shape if shape is not None else px.shape
    d = TgaDesc(w, h, pitch if pitch is not None else w * c, type_ if type_ is not None else _TGA_TYPE.get(c, -1))
    n = C.c_int(0)
    p = _L().gb200_tga_encode(px.ctypes.data + first_scanline, C.byref(d), C.byref(n))
    if not p:
        return None
    return _take_host(p, n.value).tobytes()


def tga_encode_bound(w: int, h: int, c: int) -> int:
    d = TgaDesc(w, h, w * c, _TGA_TYPE[c])
    return int(_L().gb200_tga_encode_bound(C.byref(d)))


def tga_encode_batch_device(pixels_dev: Sequence[int], shapes: Sequence[tuple], out_dev: Sequence[int], stream: int = 0):
    """gb200_tga_encode_batch_device: device pointers of gapless (h, w, c) uint8 images -> device buffers; returns the file
    lengths (0 = refused)."""
    n = len(pixels_dev)
    pin = (C.c_void_p * max(n, 1))(*pixels_dev)
    pout = (C.c_void_p * max(n, 1))(*out_dev)
    descs = (TgaDesc * max(n, 1))()
    for i, (h, w, c) in enumerate(shapes):
        descs[i] = TgaDesc(w, h, w * c, _TGA_TYPE.get(c, -1))
    lens = (C.c_int * max(n, 1))()
    _lib.check(_L().gb200_tga_encode_batch_device(n, pin, descs, pout, lens, stream), "tga_encode_batch_device")
    return [lens[i] for i in range(n)]


def bmp_encode(pixels: np.ndarray, ppmX: float = -1.0, ppmY: float = -1.0, pitch: Optional[int] = None, first_scanline: int = 0,
               shape: Optional[tuple] = None, type_: Optional[int] = None) -> Optional[bytes]:
    """saveBMP (plugins/bmp.d:166-194) of a (h, w, 3|4) uint8 image (rgb8 / rgba8): the file write_bmp writes (row padding
    zero), or None where saveBMP fails. ppmX / ppmY = Image.pixelsPerMeterX / Y (-1 = unknown)."""
    px = np.ascontiguousarray(pixels)
    h, w, c = shape if shape is not None else px.shape
    d = BmpDesc(w, h, pitch if pitch is not None else w * c, type_ if type_ is not None else {3: 9, 4: 12}.get(c, -1), ppmX, ppmY)
    n = C.c_int(0)
    p = _L().gb200_bmp_encode(px.ctypes.data + first_scanline, C.byref(d), C.byref(n))
    if not p:
        return None
    return _take_host(p, n.value).tobytes()


def bmp_load(data: bytes, req_comp: int = 0) -> Optional[PngResult]:
    """stbi_load_from_callbacks on a BMP file (stbdec.d:2263), as plugins/bmp.d:112 calls it."""
    L = _L()
    w, h, comp = C.c_int(), C.c_int(), C.c_int()
    px, py, pr = C.c_float(), C.c_float(), C.c_float()
    p = L.gb200_bmp_load(data, len(data), req_comp, C.byref(w), C.byref(h), C.byref(comp), C.byref(px), C.byref(py), C.byref(pr))
    if not p:
        return None
    c = req_comp if req_comp else comp.value
    a = _take_host(p, w.value * h.value * c)
    return PngResult(a.reshape(h.value, w.value, c), w.value, h.value, comp.value, px.value, py.value, pr.value)


def bmp_decode_batch(files: Sequence[bytes], req_comp: int = 0, files_dev: Optional[Sequence[int]] = None,
                     stream: int = 0) -> Batch:
    n, arr, lens, dev = _batch_args(files, files_dev)
    h = _L().gb200_bmp_decode_batch(n, arr, lens, dev, req_comp, stream)
    if not h:
        raise _lib.GamutB200Error("bmp_decode_batch: " + _lib.last_error())
    return Batch(h)


def identify_format(data: bytes) -> int:
    """Image.identifyFormatFromMemory (image.d:1037-1061): ImageFormat value or -1."""
    return int(_L().gb200_identify_format(data, len(data)))


def image_load(data: bytes, flags: int) -> LoadedImage:
    """gb200_image_load: decode + conversion into the requested PixelType / LayoutConstraints on the GPU, one copy back.
    The caller owns `alloc` (gb200_free)."""
    im = LoadedImage()
    _L().gb200_image_load(data, len(data), int(flags), C.byref(im))
    return im
