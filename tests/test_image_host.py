"""Host logic of gamut_b200.image (CPU only): the LoadFlags / PixelType algebra and the layout
arithmetic of Image.convertTo, checked against the reference's own unittest assertions
(internals/types.d:595-611 computeRequestedImageComponents; :165-200 layout accessors) and against the
`final switch` tables of types.d:351-602 written out by hand."""
import numpy as np

from gamut_b200 import image as im
from gamut_b200.types import *  # noqa: F401,F403
from gamut_b200.types import PixelType as P


def test_requested_components_reference_unittest():
    f = im.computeRequestedImageComponents
    assert f(LOAD_GREYSCALE) == -1
    assert f(LOAD_GREYSCALE | LOAD_NO_ALPHA) == 1
    assert f(LOAD_GREYSCALE | LOAD_ALPHA) == 2
    assert f(LOAD_GREYSCALE | LOAD_ALPHA | LOAD_NO_ALPHA) == 0
    assert f(LOAD_RGB) == -1
    assert f(LOAD_RGB | LOAD_NO_ALPHA) == 3
    assert f(LOAD_RGB | LOAD_GREYSCALE) == 0
    assert f(LOAD_RGB | LOAD_ALPHA) == 4
    assert f(LOAD_8BIT | LOAD_16BIT) == 0 and f(LOAD_PREMUL | LOAD_NO_PREMUL) == 0


def test_layout_accessors_reference_unittest():
    assert im.layoutMultiplicity(LAYOUT_MULTIPLICITY_1) == 1 and im.layoutMultiplicity(LAYOUT_MULTIPLICITY_8) == 8
    assert im.layoutTrailingPixels(LAYOUT_TRAILING_0) == 0 and im.layoutTrailingPixels(LAYOUT_TRAILING_1) == 1
    assert im.layoutTrailingPixels(LAYOUT_TRAILING_3) == 3
    assert im.layoutTrailingPixels(LAYOUT_TRAILING_7 | LAYOUT_MULTIPLICITY_8) == 7
    assert im.layoutScanlineAlignment(LAYOUT_SCANLINE_ALIGNED_1 | LAYOUT_TRAILING_7) == 1
    assert im.layoutScanlineAlignment(LAYOUT_SCANLINE_ALIGNED_128) == 128
    assert im.layoutBorderWidth(LAYOUT_BORDER_0) == 0 and im.layoutBorderWidth(LAYOUT_BORDER_1) == 1
    assert im.layoutBorderWidth(LAYOUT_BORDER_2 | LAYOUT_TRAILING_7) == 2 and im.layoutBorderWidth(LAYOUT_BORDER_3) == 3
    assert im.layoutGapless(LAYOUT_GAPLESS) and not im.layoutGapless(0)
    assert not im.layoutConstraintsValid(LAYOUT_VERT_FLIPPED | LAYOUT_VERT_STRAIGHT)
    assert not im.layoutConstraintsValid(LAYOUT_GAPLESS | LAYOUT_BORDER_1)
    assert im.layoutConstraintsValid(LAYOUT_GAPLESS | LAYOUT_VERT_STRAIGHT)


# types.d:351-602 written out as rows "type: grey rgb +alpha -alpha premul nopremul"
TABLE = """
l8 l8 rgb8 la8 l8 l8 l8
l16 l16 rgb16 la16 l16 l16 l16
lf32 lf32 rgbf32 laf32 lf32 lf32 lf32
la8 la8 rgba8 la8 l8 lap8 la8
la16 la16 rgba16 la16 l16 lap16 la16
laf32 laf32 rgbaf32 laf32 lf32 lapf32 laf32
lap8 lap8 rgbap8 lap8 l8 lap8 la8
lap16 lap16 rgbap16 lap16 l16 lap16 la16
lapf32 lapf32 rgbapf32 lapf32 lf32 lapf32 laf32
rgb8 l8 rgb8 rgba8 rgb8 rgb8 rgb8
rgb16 l16 rgb16 rgba16 rgb16 rgb16 rgb16
rgbf32 lf32 rgbf32 rgbaf32 rgbf32 rgbf32 rgbf32
rgba8 la8 rgba8 rgba8 rgb8 rgbap8 rgba8
rgba16 la16 rgba16 rgba16 rgb16 rgbap16 rgba16
rgbaf32 laf32 rgbaf32 rgbaf32 rgbf32 rgbapf32 rgbaf32
rgbap8 lap8 rgbap8 rgbap8 rgb8 rgbap8 rgba8
rgbap16 lap16 rgbap16 rgbap16 rgb16 rgbap16 rgba16
rgbapf32 lapf32 rgbapf32 rgbapf32 rgbf32 rgbapf32 rgbaf32
"""


def test_pixel_type_algebra_tables():
    fns = (im.convertPixelTypeToGreyscale, im.convertPixelTypeToRGB, im.convertPixelTypeToAddAlphaChannel,
           im.convertPixelTypeToDropAlphaChannel, im.convertPixelTypeToPremul, im.convertPixelTypeToNoPremul)
    rows = [r.split() for r in TABLE.strip().splitlines()]
    assert len(rows) == 18
    for r in rows:
        t = P[r[0]]
        for fn, exp in zip(fns, r[1:]):
            assert fn(t) == P[exp], (r[0], fn.__name__)
        base = r[0].rstrip("0123456789").removesuffix("f")
        assert im.convertPixelTypeTo8Bit(t) == P[base + "8"]
        assert im.convertPixelTypeTo16Bit(t) == P[base + "16"]
        assert im.convertPixelTypeToFP32(t) == P[base + "f32"]
    for fn in fns:
        assert fn(P.unknown) == P.unknown


def test_apply_load_flags():
    assert im.applyLoadFlags(P.rgb8, LOAD_GREYSCALE | LOAD_ALPHA | LOAD_16BIT) == P.la16
    assert im.applyLoadFlags(P.la16, LOAD_RGB | LOAD_NO_ALPHA | LOAD_FP32) == P.rgbf32
    assert im.applyLoadFlags(P.rgba8, LOAD_PREMUL) == P.rgbap8
    assert im.applyLoadFlags(P.rgbap16, LOAD_NO_PREMUL | LOAD_8BIT) == P.rgba8
    assert im.applyLoadFlags(P.rgb8, LOAD_ALPHA | LOAD_NO_ALPHA) == P.unknown
    assert im.applyLoadFlags(P.l8, 0) == P.l8


def test_allocate_pixel_storage_geometry():
    for c in (0, LAYOUT_SCANLINE_ALIGNED_64 | LAYOUT_TRAILING_7, LAYOUT_BORDER_2 | LAYOUT_MULTIPLICITY_4,
              LAYOUT_VERT_FLIPPED | LAYOUT_SCANLINE_ALIGNED_16, LAYOUT_GAPLESS, LAYOUT_BORDER_3 | LAYOUT_SCANLINE_ALIGNED_128):
        for t in (P.l8, P.rgb8, P.rgba16, P.rgbaf32):
            w, h = 13, 7
            area, off, pitch = im.allocatePixelStorage(t, w, h, c)
            px = pixelTypeSize(t)
            al = im.layoutScanlineAlignment(c)
            assert (area.ctypes.data + off) % al == 0 and abs(pitch) % al == 0
            if c & LAYOUT_VERT_FLIPPED:
                assert pitch < 0
            else:
                assert pitch > 0
            b = im.layoutBorderWidth(c)
            assert abs(pitch) >= px * (b + w + max(b, im.layoutTrailingPixels(c)))
            if al == 1 and not im.layoutTrailingPixels(c):
                assert (abs(pitch) // px - b) % im.layoutMultiplicity(c) == 0      # (border + width + right padding)
            if im.layoutGapless(c):
                assert abs(pitch) == px * w
            # every scanline (and its border rows) lies inside the allocation
            lo = min(off, off + (h - 1) * pitch) - b * abs(pitch) - b * px
            hi = max(off, off + (h - 1) * pitch) + (b + 1) * abs(pitch)
            assert lo >= 0 and hi <= area.size + abs(pitch)


def test_identify_format():
    I = im.Image
    assert I.identifyFormatFromMemory(b"\xff\xd8\xff\xe0") == ImageFormat.JPEG
    assert I.identifyFormatFromMemory(b"\x89PNG\r\n\x1a\n....") == ImageFormat.PNG
    assert I.identifyFormatFromMemory(b"qoif....") == ImageFormat.QOI
    assert I.identifyFormatFromMemory(b"qoix....") == ImageFormat.QOIX
    assert I.identifyFormatFromMemory(b"") == ImageFormat.unknown
    img = I()
    assert not img.loadFromMemory(b"garbage")
    assert img.isError() and img.errorMessage() == im.kStrImageFormatUnidentified
