"""Host-side mirror of gamut.Image for the decode/convert hot path (source/gamut/image.d).

Only the decision logic lives here -- which codec is called with which arguments, which PixelType the
buffer is adopted as, which conversion runs and into which layout. Every byte of pixel work is done by
the CUDA library through the C ABI (include/gamut_b200.h). Names and semantics follow the reference:

    Image.loadFromMemory   image.d:886-901  -> identifyFormatFromMemory (image.d:1037-1061)
                                            -> loadPNG / loadJPEG / loadQOI / loadQOIX
                                               (plugins/png.d:44-163, jpeg.d:42-104, qoi.d:48-140, qoix.d:64-146)
    Image.convertTo        image.d:1180-1332 (getAdHocLayoutConstraints :1809-1905,
                                              allocatePixelStorage internals/types.d:355-540)
    applyLoadFlags / computeRequestedImageComponents / validLoadFlags   internals/types.d:563-661
    convertPixelTypeTo*    types.d:351-602
"""
from __future__ import annotations

import numpy as np

from . import codecs
from .scanline import scanlinesConvert, scanlinesInterType
from .types import *  # noqa: F401,F403
from .types import (PixelType, ImageFormat, pixelTypeSize, LOAD_GREYSCALE, LOAD_ALPHA, LOAD_NO_ALPHA, LOAD_RGB,
                    LOAD_8BIT, LOAD_16BIT, LOAD_FP32, LOAD_PREMUL, LOAD_NO_PREMUL, LAYOUT_GAPLESS,
                    LAYOUT_VERT_FLIPPED, LAYOUT_VERT_STRAIGHT, GAMUT_MAX_IMAGE_WIDTH, GAMUT_MAX_IMAGE_HEIGHT,
                    GAMUT_UNKNOWN_ASPECT_RATIO, GAMUT_UNKNOWN_RESOLUTION)

# internals/errors.d
kStrImageDecodingFailed = "Image decoding failed"
kStrImageFormatUnidentified = "Unidentified image format"
kStrImageFormatNoLoadSupport = "Cannot decode this image format in this build"
kStrImageTooLarge = "Can't have an image that exceeds Gamut size limitations"
kStrImageWrongComponents = "Invalid number of component for image"
kStrInvalidFlags = "Invalid image decoding flags"
kStrOutOfMemory = "Out of memory"
kStrUnsupportedTypeConversion = "Unsupported image pixel type conversion"
kStrImageNotInitialized = "Uninitialized image"
kStrCannotOpenFile = "Cannot open file"                      # internals/errors.d:14
kStrFileCloseFailed = "fclose() failed"                      # internals/errors.d:15

# extensionList of the ten plugins in ImageFormat order (plugins/*.d, the `p.extensionList = ...` lines)
_EXTENSIONS = (("jpg", "jpeg", "jif", "jfif"), ("png",), ("qoi",), ("qoix",), ("dds",), ("tga",), ("gif",), ("bmp", "dib"),
               ("jxl",), ("sqz",))


def identifyImageFormatFromFilename(filename) -> "ImageFormat":
    """plugin.d:55-97: the text after the last '.' of the whole path (the whole path when it has none), compared case-
    sensitively with each plugin's extension list in ImageFormat order."""
    if filename is None:
        return ImageFormat.unknown
    name = filename if isinstance(filename, str) else bytes(filename).decode("utf-8", "surrogateescape")
    pos = len(name)
    while pos > 0 and (pos == len(name) or name[pos] != "."):
        pos -= 1
    if pos < len(name) and name[pos] == ".":
        pos += 1
    ext = name[pos:]
    for fif, exts in enumerate(_EXTENSIONS):
        if ext in exts:
            return ImageFormat(fif)
    return ImageFormat.unknown

GAMUT_MAX_IMAGE_BYTES = 0x7FFFFFFF  # internals/types.d: a gamut image allocation is bounded by int.max
LAYOUT_BORDER_MASK = 384

# ---- PixelType algebra (types.d:351-602). Types are laid out as 3 depths x 6 colour models. ----
_P = PixelType


def _model(t):   # 0 l, 1 la, 2 lap, 3 rgb, 4 rgba, 5 rgbap
    return int(t) // 3


def _depth(t):   # 0 8-bit, 1 16-bit, 2 fp32
    return int(t) % 3


def _mk(model, depth):
    return _P(model * 3 + depth)


def _map_model(t, table):
    if int(t) < 0:
        return _P.unknown
    return _mk(table[_model(t)], _depth(t))


def convertPixelTypeToGreyscale(t): return _map_model(t, (0, 1, 2, 0, 1, 2))          # types.d:351
def convertPixelTypeToRGB(t): return _map_model(t, (3, 4, 5, 3, 4, 5))                # types.d:379
def convertPixelTypeToAddAlphaChannel(t): return _map_model(t, (1, 1, 2, 4, 4, 5))    # types.d:407
def convertPixelTypeToDropAlphaChannel(t): return _map_model(t, (0, 0, 0, 3, 3, 3))   # types.d:435
def convertPixelTypeToPremul(t): return _map_model(t, (0, 2, 2, 3, 5, 5))             # types.d:463
def convertPixelTypeToNoPremul(t): return _map_model(t, (0, 1, 1, 3, 4, 4))           # types.d:491
def convertPixelTypeTo8Bit(t): return _P.unknown if int(t) < 0 else _mk(_model(t), 0)   # types.d:519
def convertPixelTypeTo16Bit(t): return _P.unknown if int(t) < 0 else _mk(_model(t), 1)  # types.d:547
def convertPixelTypeToFP32(t): return _P.unknown if int(t) < 0 else _mk(_model(t), 2)   # types.d:575


def validLoadFlags(flags: int) -> bool:  # internals/types.d:563-578
    if (flags & LOAD_GREYSCALE) and (flags & LOAD_RGB):
        return False
    if (flags & LOAD_ALPHA) and (flags & LOAD_NO_ALPHA):
        return False
    if (flags & LOAD_PREMUL) and (flags & LOAD_NO_PREMUL):
        return False
    return sum(1 for f in (LOAD_8BIT, LOAD_16BIT, LOAD_FP32) if flags & f) <= 1


def computeRequestedImageComponents(flags: int) -> int:  # internals/types.d:588-611
    if not validLoadFlags(flags):
        return 0
    if flags & LOAD_GREYSCALE:
        if flags & LOAD_ALPHA:
            return 2
        if flags & LOAD_NO_ALPHA:
            return 1
    elif flags & LOAD_RGB:
        if flags & LOAD_ALPHA:
            return 4
        if flags & LOAD_NO_ALPHA:
            return 3
    return -1


def applyLoadFlags(t, flags: int):  # internals/types.d:627-661
    if not validLoadFlags(flags):
        return _P.unknown
    for bit, fn in ((LOAD_GREYSCALE, convertPixelTypeToGreyscale), (LOAD_RGB, convertPixelTypeToRGB),
                    (LOAD_ALPHA, convertPixelTypeToAddAlphaChannel), (LOAD_NO_ALPHA, convertPixelTypeToDropAlphaChannel),
                    (LOAD_8BIT, convertPixelTypeTo8Bit), (LOAD_16BIT, convertPixelTypeTo16Bit),
                    (LOAD_FP32, convertPixelTypeToFP32), (LOAD_PREMUL, convertPixelTypeToPremul),
                    (LOAD_NO_PREMUL, convertPixelTypeToNoPremul)):
        if flags & bit:
            t = fn(t)
    return t


# ---- layout constraints (internals/types.d:165-300) ----
def layoutMultiplicity(c): return 1 << (c & 3)
def layoutTrailingPixels(c): return (1 << ((c & 0x0C) >> 2)) - 1
def layoutScanlineAlignment(c): return 1 << ((c >> 4) & 0x0F)
def layoutBorderWidth(c): return (c >> 7) & 3
def layoutGapless(c): return (c & LAYOUT_GAPLESS) != 0


def layoutConstraintsValid(c: int) -> bool:  # internals/types.d:262-283
    if (c & LAYOUT_VERT_FLIPPED) and (c & LAYOUT_VERT_STRAIGHT):
        return False
    if layoutGapless(c):
        if layoutMultiplicity(c) > 1 or layoutTrailingPixels(c) > 0 or layoutScanlineAlignment(c) > 1 \
                or layoutBorderWidth(c) > 0:
            return False
    return True


def layoutConstraintsCompatible(newer: int, older: int) -> bool:  # internals/types.d:236-259
    if (newer & LAYOUT_GAPLESS) and not (older & LAYOUT_GAPLESS):
        return False
    if (newer & LAYOUT_VERT_FLIPPED) and not (older & LAYOUT_VERT_FLIPPED):
        return False
    if (newer & LAYOUT_VERT_STRAIGHT) and not (older & LAYOUT_VERT_STRAIGHT):
        return False
    return (layoutMultiplicity(newer) <= layoutMultiplicity(older)
            and layoutTrailingPixels(newer) <= layoutTrailingPixels(older)
            and layoutScanlineAlignment(newer) <= layoutScanlineAlignment(older)
            and layoutBorderWidth(newer) <= layoutBorderWidth(older))


def _pointer_alignment(p: int) -> int:  # getPointerAlignment, internals/types.d:201-211
    for bits, flag in ((127, 112), (63, 96), (31, 80), (15, 64), (7, 48), (3, 32), (1, 16)):
        if (p & bits) == 0:
            return flag
    return 0


def imageIsValidSize(layers: int, width: int, height: int) -> bool:  # internals/types.d:150-162
    if layers < 0 or width < 0 or height < 0:
        return False
    return width <= GAMUT_MAX_IMAGE_WIDTH and height <= GAMUT_MAX_IMAGE_HEIGHT


def allocatePixelStorage(type_, width, height, constraints, bonusBytes=0):
    """allocatePixelStorage (internals/types.d:355-540) for one layer. Returns (area, data_offset, pitch) --
    `area` is the allocation (numpy bytes), `data_offset` the byte offset of the first scanline, `pitch`
    signed -- or None on failure."""
    if not imageIsValidSize(1, width, height):
        return None
    border = layoutBorderWidth(constraints)
    rowAlign = layoutScanlineAlignment(constraints)
    trailing = layoutTrailingPixels(constraints)
    mult = layoutMultiplicity(constraints)
    rightPad = (width + border + mult - 1) // mult * mult - (width + border)
    borderRight = max(border + rightPad, trailing)
    actualW = border + width + borderRight
    actualH = border + height + border
    px = pixelTypeSize(type_)
    pitch = (px * actualW + rowAlign - 1) // rowAlign * rowAlign
    need = pitch * actualH + (rowAlign - 1) + bonusBytes
    if need > GAMUT_MAX_IMAGE_BYTES:
        return None
    area = np.empty(max(need, 1), np.uint8)
    base = area.ctypes.data
    first = base + bonusBytes + pitch * border + px * border
    first = (first + rowAlign - 1) // rowAlign * rowAlign
    off = first - base
    # applyVFlipConstraintsToScanlinePointers (internals/types.d:303-320): a fresh allocation has pitch > 0
    if (constraints & LAYOUT_VERT_FLIPPED) and pitch > 0:
        if height >= 2:
            off += pitch * (height - 1)
        pitch = -pitch
    return area, off, pitch


class _MallocOwner:
    """Owns a malloc()'d pixel area returned by the C library (Image._allocArea); freed with gb200_free."""

    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            if self.ptr:
                from . import _lib
                _lib.lib().gb200_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


class Image:
    """The fields and the two hot-path methods of gamut.Image (image.d). `_area` is the owned allocation
    (numpy bytes), `_offset` the byte offset of the first scanline inside it, `_pitch` is signed."""

    def __init__(self):
        self._area = None
        self._offset = 0
        self._type = PixelType.unknown
        self._width = 0
        self._height = 0
        self._pitch = 0
        self._layoutConstraints = 0
        self._pixelAspectRatio = GAMUT_UNKNOWN_ASPECT_RATIO
        self._resolutionY = GAMUT_UNKNOWN_RESOLUTION
        self._layerCount = 0
        self._owner = None
        self._error = kStrImageNotInitialized

    # -- status (image.d:372-400)
    def isValid(self): return self._error is None
    def isError(self): return self._error is not None
    def errorMessage(self): return self._error or ""
    def error(self, msg): self._error = msg
    def hasData(self): return self._area is not None
    def type(self): return self._type
    def width(self): return self._width
    def height(self): return self._height
    def pitchInBytes(self): return self._pitch
    def layoutConstraints(self): return self._layoutConstraints
    def pixelAspectRatio(self): return self._pixelAspectRatio
    def dotsPerInchY(self): return self._resolutionY            # image.d:324

    def dotsPerInchX(self):                                     # image.d:314
        if self._resolutionY == GAMUT_UNKNOWN_RESOLUTION or self._pixelAspectRatio == GAMUT_UNKNOWN_ASPECT_RATIO:
            return GAMUT_UNKNOWN_RESOLUTION
        return float(np.float32(self._resolutionY) * np.float32(self._pixelAspectRatio))

    def pixelsPerMeterX(self):                                  # image.d:344-350 (convertMetersToInches: x * 39.37007874f)
        dpi = self.dotsPerInchX()
        return GAMUT_UNKNOWN_RESOLUTION if dpi == GAMUT_UNKNOWN_RESOLUTION else float(np.float32(dpi) * np.float32(39.37007874))

    def pixelsPerMeterY(self):                                  # image.d:355-361
        dpi = self.dotsPerInchY()
        return GAMUT_UNKNOWN_RESOLUTION if dpi == GAMUT_UNKNOWN_RESOLUTION else float(np.float32(dpi) * np.float32(39.37007874))

    def scanline(self, y: int) -> np.ndarray:
        """Bytes of scanline y (image.d scanptr)."""
        n = self._width * pixelTypeSize(self._type)
        o = self._offset + y * self._pitch
        return self._area[o:o + n]

    def pixels(self) -> np.ndarray:
        """(h, w, channels) copy in the image's component type."""
        rows = np.stack([self.scanline(y) for y in range(self._height)]) if self._height else np.zeros((0, 0), np.uint8)
        dt = (np.uint8, np.uint16, np.float32)[int(self._type) % 3]
        return rows.view(dt).reshape(self._height, self._width, -1)

    # -- identifyFormatFromMemory (image.d:1037-1061; the detect procs of all ten plugins, TGA last)
    @staticmethod
    def identifyFormatFromMemory(data: bytes) -> ImageFormat:
        return ImageFormat(codecs.identify_format(bytes(data)))

    # -- file / stream front end (image.d:859-1068, 1730-1790; io.d). The reference reads through an IOStream; here the
    #    bytes of the file or of a Python binary stream are handed to the memory path. What the front end itself decides
    #    is restated: content detection first and the file name's extension only when the content is not recognised
    #    (image.d:866-870), "Cannot open file", no load support, and -- for saving -- the format from the extension.
    @staticmethod
    def identifyFormatFromFile(path) -> ImageFormat:             # image.d:1026-1037
        try:
            with open(path, "rb") as f:
                data = f.read()
        except OSError:
            return ImageFormat.unknown
        return Image.identifyFormatFromMemory(data)

    @staticmethod
    def identifyFormatFromFileName(path) -> ImageFormat:         # image.d:1066-1069
        return identifyImageFormatFromFilename(path)

    @staticmethod
    def identifyFormatFromStream(stream) -> ImageFormat:         # image.d:1045-1061; the cursor is restored like the detect procs do
        at = stream.tell()
        data = stream.read()
        stream.seek(at)
        return Image.identifyFormatFromMemory(data)

    def _loadAs(self, fif, data: bytes, flags: int) -> bool:
        """loadFromStreamInternal (image.d:1751-1772) with the format already chosen."""
        self.__init__()
        self._error = None
        if fif == ImageFormat.unknown:
            self.error(kStrImageFormatUnidentified)
            return False
        if fif == self.identifyFormatFromMemory(data):
            return self.loadFromMemory(data, flags)              # the one-call GPU path decides the same format itself
        # the extension named a format the content does not look like: that plugin's loader still runs on the bytes
        loaders = {ImageFormat.PNG: self._loadPNG, ImageFormat.JPEG: self._loadJPEG, ImageFormat.QOI: self._loadQOI,
                   ImageFormat.QOIX: self._loadQOIX, ImageFormat.BMP: self._loadBMP, ImageFormat.TGA: self._loadTGA}
        if fif not in loaders:
            self.error(kStrImageFormatNoLoadSupport)
            return False
        loaders[fif](data, flags)
        return self.isValid()

    def loadFromFile(self, path, flags: int = 0) -> bool:         # image.d:859-873 + loadFromFileInternal (:1730-1749)
        self.__init__()
        fif = self.identifyFormatFromFile(path)
        if fif == ImageFormat.unknown:
            fif = identifyImageFormatFromFilename(path)
        try:
            with open(path, "rb") as f:
                data = f.read()
        except OSError:
            self.error(kStrCannotOpenFile)
            return False
        return self._loadAs(fif, data, flags)

    def loadFromStream(self, stream, flags: int = 0) -> bool:     # image.d:916-924: a binary file-like object, read from its cursor
        self.__init__()
        fif = self.identifyFormatFromStream(stream)
        return self._loadAs(fif, stream.read(), flags)

    def saveToStream(self, fif, stream, flags: int = 0) -> bool:  # image.d:992-1016
        if fif == ImageFormat.unknown or not self.hasData():
            return False
        data = self.saveToMemory(fif, flags)
        if data is None:                                          # no saveProc in this build, or the plugin refused the image
            return False
        return stream.write(data) == len(data)

    def saveToFile(self, path_or_fif, path=None, flags: int = 0) -> bool:
        """saveToFile(path, flags) (image.d:935-942: format from the extension) and saveToFile(fif, path, flags)
        (:953-958) -> saveToFileInternal (:1774-1786)."""
        if path is None:
            path, fif = path_or_fif, identifyImageFormatFromFilename(path_or_fif)
        else:
            fif = ImageFormat(int(path_or_fif))
        try:
            f = open(path, "wb")
        except OSError:
            return False
        with f:
            return self.saveToStream(fif, f, flags)

    def _adopt(self, px: np.ndarray, type_, pitch: int, layout: int, par: float, resY: float):
        a = np.ascontiguousarray(px).view(np.uint8).reshape(-1)
        self._area, self._offset = a, 0
        self._height, self._width = px.shape[0], px.shape[1]
        self._type, self._pitch = PixelType(type_), pitch
        self._layoutConstraints = layout
        self._pixelAspectRatio, self._resolutionY = par, resY
        self._layerCount = 1

    def loadFromMemory(self, data: bytes, flags: int = 0) -> bool:
        """image.d:886-901 + loadFromStreamInternal (:1751-1772). One call into the CUDA library: gb200_image_load
        decodes, converts into the PixelType / LayoutConstraints that `flags` ask for and copies the finished image
        back once; the allocation it returns is adopted like the plugins adopt the codec's buffer."""
        self.__init__()
        self._error = None
        r = codecs.image_load(bytes(data), flags)
        if r.error is not None or not r.alloc:
            self.error(r.error.decode() if r.error else kStrImageDecodingFailed)
            return False
        import ctypes as C
        area = np.ctypeslib.as_array((C.c_uint8 * r.alloc_bytes).from_address(r.alloc))
        self._owner = _MallocOwner(r.alloc)         # frees the area when the image lets go of it
        self._area, self._offset = area, r.data - r.alloc
        self._width, self._height, self._pitch = r.width, r.height, r.pitch
        self._type, self._layoutConstraints = PixelType(r.type), r.layout
        self._pixelAspectRatio, self._resolutionY = r.pixelAspectRatio, r.resolutionY
        self._layerCount = 1
        return True

    def loadFromMemoryStaged(self, data: bytes, flags: int = 0) -> bool:
        """The same load composed from the codec-level entry points the way the reference's plugins do it (decode to a
        gapless buffer, then convertTo): kept as the executable description of the plugin epilogues and to cross-check
        gb200_image_load in the tests."""
        self.__init__()
        self._error = None
        fif = self.identifyFormatFromMemory(data)
        if fif == ImageFormat.unknown:
            self.error(kStrImageFormatUnidentified)
            return False
        loaders = {ImageFormat.PNG: self._loadPNG, ImageFormat.JPEG: self._loadJPEG, ImageFormat.QOI: self._loadQOI,
                   ImageFormat.QOIX: self._loadQOIX, ImageFormat.BMP: self._loadBMP, ImageFormat.TGA: self._loadTGA}
        if fif not in loaders:                       # plugin.loadProc is null (image.d:1766-1770)
            self.error(kStrImageFormatNoLoadSupport)
            return False
        loaders[fif](data, flags)
        return self.isValid()

    def _loadTGA(self, data, flags):  # plugins/tga.d:45-105
        px = codecs.tga_load(data)
        if px is None:
            return self.error(kStrImageDecodingFailed)
        comps = px.shape[2]
        t = (None, _P.l8, _P.la8, _P.rgb8, _P.rgba8)[comps]
        self._adopt(px, t, px.shape[1] * comps, 0, GAMUT_UNKNOWN_ASPECT_RATIO, GAMUT_UNKNOWN_RESOLUTION)
        self.convertTo(applyLoadFlags(self._type, flags), flags & 0xFFFF)

    def _loadBMP(self, data, flags):  # plugins/bmp.d:93-163
        req = computeRequestedImageComponents(flags)
        if req == 0:
            return self.error(kStrInvalidFlags)
        if req == -1:
            req = 0
        r = codecs.bmp_load(data, req)
        if r is None:
            return self.error(kStrImageDecodingFailed)
        comps = req if req else r.file_channels
        if not imageIsValidSize(1, r.width, r.height):
            return self.error(kStrImageTooLarge)
        t = (None, _P.l8, _P.la8, _P.rgb8, _P.rgba8)[comps]
        par = GAMUT_UNKNOWN_ASPECT_RATIO if r.pixelRatio == -1 else r.pixelRatio
        resY = GAMUT_UNKNOWN_RESOLUTION if r.ppmY == -1 else float(np.float32(r.ppmY) / np.float32(39.37007874))  # convertInchesToMeters
        self._adopt(r.pixels, t, r.width * comps, 0, par, resY)
        self.convertTo(applyLoadFlags(self._type, flags), flags & 0xFFFF)

    def _loadPNG(self, data, flags):  # plugins/png.d:44-163
        is16 = codecs.png_is16(data)
        req = computeRequestedImageComponents(flags)
        if req == 0:
            return self.error(kStrInvalidFlags)
        if req == -1:
            req = 0
        to16 = is16
        if flags & LOAD_8BIT:
            to16 = False
        if flags & LOAD_16BIT:
            to16 = True
        r = codecs.png_load(data, req, to16)
        if r is None:
            return self.error(kStrImageDecodingFailed)
        comps = req if req else r.file_channels
        if not imageIsValidSize(1, r.width, r.height):
            return self.error(kStrImageTooLarge)
        t = (None, _P.l8, _P.la8, _P.rgb8, _P.rgba8)[comps] if not to16 else (None, _P.l16, _P.la16, _P.rgb16, _P.rgba16)[comps]
        par = GAMUT_UNKNOWN_ASPECT_RATIO if r.pixelRatio == -1 else r.pixelRatio
        resY = GAMUT_UNKNOWN_RESOLUTION if r.ppmY == -1 else float(np.float32(r.ppmY) / np.float32(39.37007874))  # convertInchesToMeters, types.d:127
        self._adopt(r.pixels, t, r.width * comps * (2 if to16 else 1), 0, par, resY)
        self.convertTo(applyLoadFlags(self._type, flags), flags & 0xFFFF)

    def _loadJPEG(self, data, flags):  # plugins/jpeg.d:42-104
        req = computeRequestedImageComponents(flags)
        if req == 0:
            return self.error(kStrInvalidFlags)
        if req == 2:
            req = -1
        r = codecs.jpeg_load(data, req)
        if r is None:
            # a valid non-interleaved multi-scan sequential file is a kind this build cannot decode, not a broken one
            return self.error(kStrImageFormatNoLoadSupport if codecs.jpeg_probe(data) > 0 else kStrImageDecodingFailed)
        if r.actual_comps not in (1, 3, 4):
            return self.error(kStrImageWrongComponents)
        if not imageIsValidSize(1, r.width, r.height):
            return self.error(kStrImageTooLarge)
        comps = r.actual_comps if req == -1 else req
        t = {1: _P.l8, 3: _P.rgb8, 4: _P.rgba8}[comps]
        par = GAMUT_UNKNOWN_ASPECT_RATIO if r.pixelAspectRatio == -1 else r.pixelAspectRatio
        resY = GAMUT_UNKNOWN_RESOLUTION if r.dotsPerInchY == -1 else r.dotsPerInchY
        self._adopt(r.pixels, t, r.width * comps, 0, par, resY)
        self.convertTo(applyLoadFlags(self._type, flags), flags & 0xFFFF)

    def _loadQOI(self, data, flags):  # plugins/qoi.d:48-140
        req = computeRequestedImageComponents(flags)
        if req == 0:
            return self.error(kStrInvalidFlags)
        if req in (-1, 1, 2):
            req = 0
        r = codecs.qoi_decode(data, req)
        if r is None:
            return self.error(kStrImageDecodingFailed)
        px, desc = r
        if not imageIsValidSize(1, desc.width, desc.height):
            return self.error(kStrImageTooLarge)
        comps = desc.channels if req == 0 else req
        t = {3: _P.rgb8, 4: _P.rgba8}[comps]
        # the reference sets _pitch = desc.channels * width (plugins/qoi.d:131), which equals the buffer's
        # real pitch only when the requested channel count equals the file's; the buffer itself is gapless
        self._adopt(px, t, comps * desc.width, 0, GAMUT_UNKNOWN_ASPECT_RATIO, GAMUT_UNKNOWN_RESOLUTION)
        self.convertTo(applyLoadFlags(self._type, flags), flags & 0xFFFF)

    def _loadQOIX(self, data, flags):  # plugins/qoix.d:64-146
        req = computeRequestedImageComponents(flags)
        if req == 0:
            return self.error(kStrInvalidFlags)
        r = codecs.qoix_decode(data, flags)
        if r is None:
            return self.error(kStrImageDecodingFailed)
        px, desc, t = r
        if not imageIsValidSize(1, desc.width, desc.height):
            return self.error(kStrImageTooLarge)
        self._adopt(px, t, desc.pitchBytes, 0, desc.pixelAspectRatio, desc.resolutionY)
        self.convertTo(applyLoadFlags(self._type, flags), flags & 0xFFFF)

    # -- getAdHocLayoutConstraints (image.d:1809-1905)
    # -- Image.saveToMemory (image.d:966) for the save paths that are built: saveQOIX (plugins/qoix.d:156-241) of a
    #    greyscale image (10-bit -> qoiplane10_encode, 8-bit -> qoiplane_encode) or of an 8-bit RGB(A) image (-> qoix_encode,
    #    QOI2AVG), saveTGA (plugins/tga.d:123-149) of an
    #    l8 / la8 / rgb8 / rgba8 image, saveBMP (plugins/bmp.d:166-194) of an rgb8 / rgba8 image, and saveQOI (plugins/qoi.d:150-185) of
    #    an rgb8 / rgba8 image. Returns the file bytes or None (the reference returns a null slice when the plugin's
    #    saveProc fails or the format has none).
    def saveToMemory(self, fmt, flags: int = 0):
        if not self.hasData():
            return None
        import ctypes as C
        t = PixelType(int(self._type))
        first = self._area.ctypes.data + self._offset
        if int(fmt) == int(ImageFormat.QOI):
            if t not in (PixelType.rgb8, PixelType.rgba8):
                return None                                # saveQOI: "not supported" (plugins/qoi.d:165-169)
            d = codecs.QoiDesc(self._width, self._height, 3 if t == PixelType.rgb8 else 4, 0)     # QOI_SRGB (:159)
            n = C.c_int(0)
            p = codecs._L().gb200_qoi_encode(first, C.byref(d), self._pitch, C.byref(n))
            return codecs._take_host(p, n.value).tobytes() if p else None
        if int(fmt) == int(ImageFormat.BMP):
            if t not in (PixelType.rgb8, PixelType.rgba8):
                return None                                # saveBMP: "can save RGB and RGBA 8-bit images" (plugins/bmp.d:173-183)
            d = codecs.BmpDesc(self._width, self._height, self._pitch, int(t), self.pixelsPerMeterX(), self.pixelsPerMeterY())
            n = C.c_int(0)
            p = codecs._L().gb200_bmp_encode(first, C.byref(d), C.byref(n))
            return codecs._take_host(p, n.value).tobytes() if p else None
        if int(fmt) == int(ImageFormat.TGA):
            if t not in (PixelType.l8, PixelType.la8, PixelType.rgb8, PixelType.rgba8):
                return None                                # TGAEncoder.initialize: unsupported format (codecs/tga.d:108-110)
            d = codecs.TgaDesc(self._width, self._height, self._pitch, int(t))
            n = C.c_int(0)
            p = codecs._L().gb200_tga_encode(first, C.byref(d), C.byref(n))
            return codecs._take_host(p, n.value).tobytes() if p else None
        if int(fmt) != int(ImageFormat.QOIX):
            return None
        if t in (PixelType.l16, PixelType.la16, PixelType.lap16):
            channels, bitdepth = (1 if t == PixelType.l16 else 2), 10          # -> qoiplane10_encode (plugins/qoix.d:199-212)
        elif t in (PixelType.l8, PixelType.la8, PixelType.lap8):
            channels, bitdepth = (1 if t == PixelType.l8 else 2), 8            # -> qoiplane_encode (plugins/qoix.d:172-184)
        elif t in (PixelType.rgb8, PixelType.rgba8, PixelType.rgbap8):
            channels, bitdepth = (3 if t == PixelType.rgb8 else 4), 8          # -> qoix_encode, QOI2AVG (plugins/qoix.d:185-198)
        elif t in (PixelType.rgb16, PixelType.rgba16, PixelType.rgbap16):
            channels, bitdepth = (3 if t == PixelType.rgb16 else 4), 10        # -> qoi10b_encode (plugins/qoix.d:213-228)
        else:
            return None                                    # fp32 types: saveQOIX "not supported" (plugins/qoix.d:229-230)
        if self._pitch < self._width * pixelTypeSize(t):
            return None                                    # vertically flipped storage: not taken by the C entry point
        d = codecs.QoixDesc(self._width, self._height, self._pitch, channels, bitdepth,
                            2 if t in (PixelType.lap16, PixelType.lap8, PixelType.rgbap8, PixelType.rgbap16) else 0, 0, self._pixelAspectRatio, self._resolutionY)
        n = C.c_int(0)
        p = codecs._L().gb200_qoix_encode(first, C.byref(d), C.byref(n))
        return codecs._take_host(p, n.value).tobytes() if p else None

    def getAdHocLayoutConstraints(self) -> int:
        pitch = self._pitch
        absPitch = abs(pitch)
        px = pixelTypeSize(self._type)
        excess = (absPitch - self._width * px) // px
        c = 0
        multi = layoutMultiplicity(self._layoutConstraints)
        gap = 8 if excess >= 7 else 4 if excess >= 3 else 2 if excess >= 1 else 1
        wdiv = 8 if self._width % 8 == 0 else 4 if self._width % 4 == 0 else 2 if self._width % 2 == 0 else 1
        multi = max(multi, gap, wdiv)
        c |= {1: 0, 2: 1, 4: 2, 8: 3}[multi]
        c |= 12 if excess >= 7 else 8 if excess >= 3 else 4 if excess >= 1 else 0
        c |= min(_pointer_alignment(self._area.ctypes.data + self._offset), _pointer_alignment(absPitch))
        if pitch >= 0:
            c |= LAYOUT_VERT_STRAIGHT
        if pitch <= 0:
            c |= LAYOUT_VERT_FLIPPED
        if pitch == absPitch:            # sic: the reference infers "gapless" from the sign of the pitch (image.d:1886)
            c |= LAYOUT_GAPLESS
        c |= self._layoutConstraints & LAYOUT_BORDER_MASK
        return c

    def convertTo(self, targetType, layoutConstraints: int = 0) -> bool:
        """image.d:1180-1332. The scanline conversion itself is gb200_scanlines_convert (CUDA)."""
        if int(targetType) == PixelType.unknown:
            self.error(kStrUnsupportedTypeConversion)
            return False
        assert layoutConstraintsValid(layoutConstraints)
        if not self.hasData():
            self._type = PixelType(targetType)
            self._layoutConstraints = layoutConstraints
            return True
        compatible = layoutConstraintsCompatible(layoutConstraints, self.getAdHocLayoutConstraints())
        if self._type == targetType and compatible:
            self._layoutConstraints = layoutConstraints
            return True
        if (self._width == 0 or self._height == 0) and compatible:
            self._layoutConstraints = layoutConstraints
            return True
        # image.d:1233-1236: the allocation also holds one intermediate scanline when the type changes
        bonus = self._width * pixelTypeSize(scanlinesInterType(self._type, targetType)) if targetType != self._type else 0
        st = allocatePixelStorage(targetType, self._width, self._height, layoutConstraints, bonus)
        if st is None:
            self.error(kStrOutOfMemory)
            return False
        area, off, pitch = st
        ok = scanlinesConvert(self._type, self._area, self._pitch, targetType, area, pitch, self._width, self._height,
                              src_offset=self._offset, dest_offset=off)
        if not ok:
            self.error(kStrUnsupportedTypeConversion)
            return False
        self._area, self._offset, self._pitch = area, off, pitch
        self._type = PixelType(targetType)
        self._layoutConstraints = layoutConstraints
        self._error = None
        return True

    # -- the convertTo family (image.d:1086-1172): each is convertTo(<type algebra>(_type), layoutConstraints)
    def setLayout(self, layoutConstraints: int) -> bool: return self.convertTo(self._type, layoutConstraints)
    def convertToGreyscale(self, lc: int = 0) -> bool: return self.convertTo(convertPixelTypeToGreyscale(self._type), lc)
    def convertToGreyscaleAlpha(self, lc: int = 0) -> bool:
        return self.convertTo(convertPixelTypeToAddAlphaChannel(convertPixelTypeToGreyscale(self._type)), lc)
    def convertToRGB(self, lc: int = 0) -> bool: return self.convertTo(convertPixelTypeToRGB(self._type), lc)
    def convertToRGBA(self, lc: int = 0) -> bool:
        return self.convertTo(convertPixelTypeToAddAlphaChannel(convertPixelTypeToRGB(self._type)), lc)
    def addAlphaChannel(self, lc: int = 0) -> bool: return self.convertTo(convertPixelTypeToAddAlphaChannel(self._type), lc)
    def dropAlphaChannel(self, lc: int = 0) -> bool: return self.convertTo(convertPixelTypeToDropAlphaChannel(self._type), lc)
    def premultiply(self, lc: int = 0) -> bool: return self.convertTo(convertPixelTypeToPremul(self._type), lc)
    def unpremultiply(self, lc: int = 0) -> bool: return self.convertTo(convertPixelTypeToNoPremul(self._type), lc)
    def convertTo8Bit(self, lc: int = 0) -> bool: return self.convertTo(convertPixelTypeTo8Bit(self._type), lc)
    def convertTo16Bit(self, lc: int = 0) -> bool: return self.convertTo(convertPixelTypeTo16Bit(self._type), lc)
    def convertToFP32(self, lc: int = 0) -> bool: return self.convertTo(convertPixelTypeToFP32(self._type), lc)
