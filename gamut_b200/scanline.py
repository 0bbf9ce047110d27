"""Host-side mirror of the reference's public scanline API (source/gamut/scanline.d:21-121),
calling the CUDA C-ABI. Same names, argument meaning and return convention (bool)."""
from __future__ import annotations

import numpy as np

from . import _lib
from .types import PixelType, pixelTypeIs8Bit, pixelTypeSize


def scanlinesInterType(srcType: int, destType: int) -> PixelType:
    """scanline.d:25-31."""
    if pixelTypeIs8Bit(srcType) and pixelTypeIs8Bit(destType):
        return PixelType.rgba8
    return PixelType.rgbaf32


def _addr(buf, off: int) -> int:
    if isinstance(buf, np.ndarray):
        return buf.ctypes.data + off
    return int(buf) + off  # raw address


def scanlinesConvert(srcType: int, src, srcPitch: int, destType: int, dest, destPitch: int,
                     width: int, height: int, interType=None, interBuf=None,
                     src_offset: int = 0, dest_offset: int = 0) -> bool:
    """scanlinesConvert (scanline.d:70). `src`/`dest` are numpy byte buffers (or raw host addresses);
    `*_offset` is the byte offset of the first scanline (needed for negative pitches).
    interType/interBuf are accepted for signature parity and ignored (stages are fused on the GPU)."""
    return bool(_lib.lib().gb200_scanlines_convert(int(srcType), _addr(src, src_offset), int(srcPitch),
                                                   int(destType), _addr(dest, dest_offset), int(destPitch),
                                                   int(width), int(height)))


def scanlinesCopy(type_: int, src, srcPitch: int, dest, destPitch: int, width: int, height: int,
                  src_offset: int = 0, dest_offset: int = 0) -> bool:
    """scanlinesCopy (scanline.d:37)."""
    return scanlinesConvert(type_, src, srcPitch, type_, dest, destPitch, width, height,
                            src_offset=src_offset, dest_offset=dest_offset)


def scanlinesConvertDevice(srcType: int, src_ptr: int, srcPitch: int, destType: int, dest_ptr: int,
                           destPitch: int, width: int, height: int, stream: int = 0) -> bool:
    """Device-resident variant: raw device addresses, asynchronous on `stream` (a cudaStream_t)."""
    return bool(_lib.lib().gb200_scanlines_convert_device(int(srcType), src_ptr, int(srcPitch), int(destType),
                                                          dest_ptr, int(destPitch), int(width), int(height),
                                                          stream))
