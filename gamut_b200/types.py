"""PixelType / LoadFlags / LayoutConstraints -- integer values equal the reference's
(source/gamut/types.d:32-59, 141-197, 271-348); they are ABI for the C shim."""
from __future__ import annotations

import enum


class PixelType(enum.IntEnum):  # types.d:32-59
    unknown = -1
    l8 = 0
    l16 = 1
    lf32 = 2
    la8 = 3
    la16 = 4
    laf32 = 5
    lap8 = 6
    lap16 = 7
    lapf32 = 8
    rgb8 = 9
    rgb16 = 10
    rgbf32 = 11
    rgba8 = 12
    rgba16 = 13
    rgbaf32 = 14
    rgbap8 = 15
    rgbap16 = 16
    rgbapf32 = 17


_SIZES = (1, 2, 4, 2, 4, 8, 2, 4, 8, 3, 6, 12, 4, 8, 16, 4, 8, 16)
_CHANNELS = (1, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3, 3, 4, 4, 4, 4, 4, 4)


def pixelTypeSize(t: int) -> int:  # types.d:62-86
    assert 0 <= int(t) <= 17
    return _SIZES[int(t)]


def pixelTypeNumChannels(t: int) -> int:
    return _CHANNELS[int(t)]


def pixelTypeComponentSize(t: int) -> int:
    return (1, 2, 4)[int(t) % 3]


def pixelTypeIs8Bit(t: int) -> bool:  # internals/types.d:99-111
    return int(t) in (PixelType.l8, PixelType.la8, PixelType.rgb8, PixelType.rgba8)


class ImageFormat(enum.IntEnum):  # types.d:14-28
    unknown = -1
    JPEG = 0
    PNG = 1
    QOI = 2
    QOIX = 3
    DDS = 4
    TGA = 5
    GIF = 6
    BMP = 7
    JXL = 8
    SQZ = 9


# LoadFlags, types.d:141-197
LOAD_NORMAL = 0
LOAD_GREYSCALE = 0x1_0000
LOAD_ALPHA = 0x2_0000
LOAD_NO_ALPHA = 0x4_0000
LOAD_RGB = 0x8_0000
LOAD_8BIT = 0x10_0000
LOAD_16BIT = 0x20_0000
LOAD_FP32 = 0x40_0000
LOAD_NO_PIXELS = 0x80_0000
LOAD_PREMUL = 0x100_0000
LOAD_NO_PREMUL = 0x200_0000

GAMUT_UNKNOWN_RESOLUTION = -1
GAMUT_UNKNOWN_ASPECT_RATIO = -1
GAMUT_MAX_IMAGE_WIDTH = 16777216
GAMUT_MAX_IMAGE_HEIGHT = 16777216

# LayoutConstraints, types.d:271-348
LAYOUT_DEFAULT = 0
LAYOUT_MULTIPLICITY_1 = 0
LAYOUT_MULTIPLICITY_2 = 1
LAYOUT_MULTIPLICITY_4 = 2
LAYOUT_MULTIPLICITY_8 = 3
LAYOUT_TRAILING_0 = 0
LAYOUT_TRAILING_1 = 4
LAYOUT_TRAILING_3 = 8
LAYOUT_TRAILING_7 = 12
LAYOUT_SCANLINE_ALIGNED_1 = 0
LAYOUT_SCANLINE_ALIGNED_2 = 16
LAYOUT_SCANLINE_ALIGNED_4 = 32
LAYOUT_SCANLINE_ALIGNED_8 = 48
LAYOUT_SCANLINE_ALIGNED_16 = 64
LAYOUT_SCANLINE_ALIGNED_32 = 80
LAYOUT_SCANLINE_ALIGNED_64 = 96
LAYOUT_SCANLINE_ALIGNED_128 = 112
LAYOUT_BORDER_0 = 0
LAYOUT_BORDER_1 = 128
LAYOUT_BORDER_2 = 256
LAYOUT_BORDER_3 = 384
LAYOUT_VERT_FLIPPED = 512
LAYOUT_VERT_STRAIGHT = 1024
LAYOUT_GAPLESS = 2048
