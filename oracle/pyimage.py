"""CPU oracle of Image.loadFromMemory / Image.convertTo (TEST INFRASTRUCTURE ONLY).

A second, independent restatement of the reference's plugin epilogues and conversion decision logic, written
straight from the D text and composed ONLY of oracle pieces (oracle/*.c through pyoracle): the product's host mirror
(gamut_b200/image.py) is never imported here. tests/test_image_gpu.py compares the two on the reference's own
scenarios (examples/test-suite/source/main.d, image.d:2112-2183).

    loadPNG   plugins/png.d:44-163      loadJPEG  plugins/jpeg.d:42-104
    loadQOI   plugins/qoi.d:48-140      loadQOIX  plugins/qoix.d:64-146
    convertTo image.d:1180-1332         getAdHocLayoutConstraints image.d:1809-1905
    allocatePixelStorage internals/types.d:355-540   applyVFlipConstraintsToScanlinePointers :303-320
    applyLoadFlags / computeRequestedImageComponents / validLoadFlags internals/types.d:563-661
"""
from __future__ import annotations

import numpy as np

from . import pyoracle

# types.d:32-59 (integers are ABI)
L8, L16, LF32, LA8, LA16, LAF32, LAP8, LAP16, LAPF32, RGB8, RGB16, RGBF32, RGBA8, RGBA16, RGBAF32, RGBAP8, RGBAP16, RGBAPF32 = range(18)
SIZE = (1, 2, 4, 2, 4, 8, 2, 4, 8, 3, 6, 12, 4, 8, 16, 4, 8, 16)            # pixelTypeSize, types.d:62-86
# LoadFlags types.d:141-197
GREY, ALPHA, NO_ALPHA, RGB, B8, B16, FP32, PREMUL, NO_PREMUL = (0x10000, 0x20000, 0x40000, 0x80000, 0x100000, 0x200000,
                                                               0x400000, 0x1000000, 0x2000000)
VERT_FLIPPED, VERT_STRAIGHT, GAPLESS, BORDER_MASK = 512, 1024, 2048, 384

E_DECODE = "Image decoding failed"
E_UNIDENT = "Unidentified image format"
E_NOLOAD = "Cannot decode this image format in this build"
E_FLAGS = "Invalid image decoding flags"
E_COMPONENTS = "Invalid number of component for image"
E_CONV = "Unsupported image pixel type conversion"
E_OOM = "Out of memory"


def valid_load_flags(f):                                   # internals/types.d:563-578
    if (f & GREY) and (f & RGB):
        return False
    if (f & ALPHA) and (f & NO_ALPHA):
        return False
    if (f & PREMUL) and (f & NO_PREMUL):
        return False
    return sum(1 for b in (B8, B16, FP32) if f & b) <= 1


def requested_components(f):                               # internals/types.d:588-611
    if not valid_load_flags(f):
        return 0
    if f & GREY:
        if f & ALPHA:
            return 2
        if f & NO_ALPHA:
            return 1
    elif f & RGB:
        if f & ALPHA:
            return 4
        if f & NO_ALPHA:
            return 3
    return -1


# the `final switch` tables of types.d:351-602, spelled out per type
TO_GREY = (L8, L16, LF32, LA8, LA16, LAF32, LAP8, LAP16, LAPF32, L8, L16, LF32, LA8, LA16, LAF32, LAP8, LAP16, LAPF32)
TO_RGB = (RGB8, RGB16, RGBF32, RGBA8, RGBA16, RGBAF32, RGBAP8, RGBAP16, RGBAPF32, RGB8, RGB16, RGBF32, RGBA8, RGBA16, RGBAF32,
          RGBAP8, RGBAP16, RGBAPF32)
ADD_A = (LA8, LA16, LAF32, LA8, LA16, LAF32, LAP8, LAP16, LAPF32, RGBA8, RGBA16, RGBAF32, RGBA8, RGBA16, RGBAF32, RGBAP8,
         RGBAP16, RGBAPF32)
DROP_A = (L8, L16, LF32, L8, L16, LF32, L8, L16, LF32, RGB8, RGB16, RGBF32, RGB8, RGB16, RGBF32, RGB8, RGB16, RGBF32)
TO_PREMUL = (L8, L16, LF32, LAP8, LAP16, LAPF32, LAP8, LAP16, LAPF32, RGB8, RGB16, RGBF32, RGBAP8, RGBAP16, RGBAPF32, RGBAP8,
             RGBAP16, RGBAPF32)
TO_NOPREMUL = (L8, L16, LF32, LA8, LA16, LAF32, LA8, LA16, LAF32, RGB8, RGB16, RGBF32, RGBA8, RGBA16, RGBAF32, RGBA8, RGBA16,
               RGBAF32)


def apply_load_flags(t, f):                                # internals/types.d:627-661 (order of the tests matters)
    if not valid_load_flags(f):
        return -1
    if f & GREY:
        t = TO_GREY[t]
    if f & RGB:
        t = TO_RGB[t]
    if f & ALPHA:
        t = ADD_A[t]
    if f & NO_ALPHA:
        t = DROP_A[t]
    if f & B8:
        t = t - t % 3
    if f & B16:
        t = t - t % 3 + 1
    if f & FP32:
        t = t - t % 3 + 2
    if f & PREMUL:
        t = TO_PREMUL[t]
    if f & NO_PREMUL:
        t = TO_NOPREMUL[t]
    return t


def _layout(c):
    # internals/types.d:163-200. layoutScanlineAlignment masks with 0x0f, which includes bit 7 (the low border bit):
    # LAYOUT_BORDER_1 / _3 therefore also request 256-byte scanline alignment in the reference -- restated, not fixed
    return dict(mult=1 << (c & 3), trailing=(1 << ((c >> 2) & 3)) - 1, align=1 << ((c >> 4) & 0x0F), border=(c >> 7) & 3,
                gapless=bool(c & GAPLESS), flipped=bool(c & VERT_FLIPPED), straight=bool(c & VERT_STRAIGHT))


def _ptr_align_flag(p):                                    # getPointerAlignment, internals/types.d:201-211
    for bits, flag in ((127, 112), (63, 96), (31, 80), (15, 64), (7, 48), (3, 32), (1, 16)):
        if (p & bits) == 0:
            return flag
    return 0


class OImage:
    """State of a gamut.Image after a load: area (numpy bytes), offset of scanline 0, signed pitch."""

    def __init__(self):
        self.area, self.off, self.pitch = None, 0, 0
        self.type, self.w, self.h, self.layout = -1, 0, 0, 0
        self.par, self.resY = -1.0, -1.0
        self.error = None

    def scanline(self, y):
        n = self.w * SIZE[self.type]
        o = self.off + y * self.pitch
        return self.area[o:o + n]

    def rows(self):
        return np.stack([self.scanline(y) for y in range(self.h)]) if self.h else np.zeros((0, 0), np.uint8)

    # image.d:1809-1905
    def adhoc(self):
        ap = abs(self.pitch)
        px = SIZE[self.type]
        excess = (ap - self.w * px) // px
        c = 0
        multi = 1 << (self.layout & 3)
        gap = 8 if excess >= 7 else 4 if excess >= 3 else 2 if excess >= 1 else 1
        wd = 1
        for m in (2, 4, 8):
            if self.w % m == 0:
                wd = m
        multi = max(multi, gap, wd)
        c |= {1: 0, 2: 1, 4: 2, 8: 3}[multi]
        c |= 12 if excess >= 7 else 8 if excess >= 3 else 4 if excess >= 1 else 0
        c |= min(_ptr_align_flag(self.area.ctypes.data + self.off), _ptr_align_flag(ap))
        if self.pitch >= 0:
            c |= VERT_STRAIGHT
        if self.pitch <= 0:
            c |= VERT_FLIPPED
        if self.pitch == ap:                               # image.d:1886 (gapless inferred from the pitch sign; one layer)
            c |= GAPLESS
        return c | (self.layout & BORDER_MASK)

    # image.d:1180-1332
    def convert_to(self, target, layout):
        if target == -1:
            self.error = E_CONV
            return False
        if self.area is None:
            self.type, self.layout = target, layout
            return True
        new, old = _layout(layout), _layout(self.adhoc())
        compatible = not ((new["gapless"] and not old["gapless"]) or (new["flipped"] and not old["flipped"])
                          or (new["straight"] and not old["straight"]) or new["mult"] > old["mult"]
                          or new["trailing"] > old["trailing"] or new["align"] > old["align"]
                          or new["border"] > old["border"])      # layoutConstraintsCompatible, internals/types.d:236-259
        if (self.type == target or self.w == 0 or self.h == 0) and compatible:
            self.layout = layout
            return True
        # allocatePixelStorage, internals/types.d:355-540 (one layer)
        border, align, trailing, mult = new["border"], new["align"], new["trailing"], new["mult"]
        right_pad = (self.w + border + mult - 1) // mult * mult - (self.w + border)
        border_right = max(border + right_pad, trailing)
        actual_w, actual_h = border + self.w + border_right, border + self.h + border
        px = SIZE[target]
        pitch = (px * actual_w + align - 1) // align * align
        inter = pyoracle.lib().or_scanlinesInterType(self.type, target)
        bonus = self.w * SIZE[inter] if target != self.type else 0            # image.d:1233-1236
        need = pitch * actual_h + (align - 1) + bonus
        if need > 0x7FFFFFFF:
            self.error = E_OOM
            return False
        area = np.zeros(max(need, 1), np.uint8)
        first = area.ctypes.data + bonus + pitch * border + px * border
        first = (first + align - 1) // align * align
        off = first - area.ctypes.data
        if new["flipped"] and pitch > 0:                   # applyVFlipConstraintsToScanlinePointers :303-320
            if self.h >= 2:
                off += pitch * (self.h - 1)
            pitch = -pitch
        if not pyoracle.scanlines_convert(self.type, self.area, self.pitch, target, area, pitch, self.w, self.h,
                                          src_off=self.off, dst_off=off):
            self.error = E_CONV
            return False
        self.area, self.off, self.pitch, self.type, self.layout, self.error = area, off, pitch, target, layout, None
        return True


def _adopt(im, px, type_, pitch, par, resY):
    im.area = np.ascontiguousarray(px).view(np.uint8).reshape(-1)
    im.off, im.pitch, im.type = 0, pitch, type_
    im.h, im.w = px.shape[0], px.shape[1]
    im.layout, im.par, im.resY = 0, par, resY


def load_from_memory(data: bytes, flags: int = 0) -> OImage:
    """Image.loadFromMemory (image.d:886-901) -> detect (plugins/*.d detectProc) -> loadProc -> convertTo."""
    im = OImage()
    req = requested_components(flags)
    if data[:2] == b"\xff\xd8":                            # plugins/jpeg.d:42-104
        if req == 0:
            im.error = E_FLAGS
            return im
        if req == 2:
            req = -1
        r = pyoracle.jpeg_load(data, req)
        if r is None:
            im.error = E_DECODE
            return im
        px, actual, par, dpi = r
        if actual not in (1, 3, 4):
            im.error = E_COMPONENTS
            return im
        comps = actual if req == -1 else req
        _adopt(im, px, {1: L8, 3: RGB8, 4: RGBA8}[comps], px.shape[1] * comps, -1.0 if par == -1 else par, -1.0 if dpi == -1 else dpi)
    elif data[:8] == b"\x89PNG\r\n\x1a\n":                # plugins/png.d:44-163
        if req == 0:
            im.error = E_FLAGS
            return im
        if req == -1:
            req = 0
        to16 = pyoracle.png_is16(data)
        if flags & B8:
            to16 = False
        if flags & B16:
            to16 = True
        px, info = pyoracle.png_load(data, req, 1 if to16 else 0)
        if px is None:
            im.error = E_DECODE
            return im
        comps = req if req else info.file_channels
        t = (None, L16, LA16, RGB16, RGBA16)[comps] if to16 else (None, L8, LA8, RGB8, RGBA8)[comps]
        resY = -1.0 if info.ppmY == -1 else float(np.float32(info.ppmY) / np.float32(39.37007874))
        _adopt(im, px, t, info.width * comps * (2 if to16 else 1), -1.0 if info.pixelRatio == -1 else info.pixelRatio, resY)
    elif data[:4] == b"qoif":                              # plugins/qoi.d:48-140
        if req == 0:
            im.error = E_FLAGS
            return im
        if req in (-1, 1, 2):
            req = 0
        r = pyoracle.qoi_decode(data, req)
        if r is None:
            im.error = E_DECODE
            return im
        px, desc = r
        comps = desc.channels if req == 0 else req
        _adopt(im, px, {3: RGB8, 4: RGBA8}[comps], comps * desc.width, -1.0, -1.0)
    elif data[:4] == b"qoix":                              # plugins/qoix.d:64-146
        if req == 0:
            im.error = E_FLAGS
            return im
        r = pyoracle.qoix_decode(data, flags)
        if r is None:
            im.error = E_DECODE
            return im
        px, desc, t = r
        _adopt(im, px, t, desc.pitchBytes, desc.pixelAspectRatio, desc.resolutionY)
    elif pyoracle.identify_format(data) == 7:              # plugins/bmp.d:93-163
        if req == 0:
            im.error = E_FLAGS
            return im
        if req == -1:
            req = 0
        r = pyoracle.bmp_load(data, req)
        if r is None:
            im.error = E_DECODE
            return im
        px, comp, ppmX, ppmY, ratio = r
        comps = req if req else comp
        resY = -1.0 if ppmY == -1 else float(np.float32(ppmY) / np.float32(39.37007874))
        _adopt(im, px, (None, L8, LA8, RGB8, RGBA8)[comps], px.shape[1] * comps, -1.0 if ratio == -1 else ratio, resY)
    elif pyoracle.identify_format(data) == 5:              # plugins/tga.d:45-105 (no component-flag logic of its own)
        px = pyoracle.tga_load(data)
        if px is None:
            im.error = E_DECODE
            return im
        comps = px.shape[2]
        _adopt(im, px, (None, L8, LA8, RGB8, RGBA8)[comps], px.shape[1] * comps, -1.0, -1.0)
    elif pyoracle.identify_format(data) >= 0:              # detected, but the build has no loader for it (image.d:1766-1770)
        im.error = E_NOLOAD
        return im
    else:
        im.error = E_UNIDENT
        return im
    im.convert_to(apply_load_flags(im.type, flags), flags & 0xFFFF)
    return im
