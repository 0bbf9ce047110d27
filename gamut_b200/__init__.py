"""gamut_b200 -- Blackwell-native (sm_100a) decode/convert engine behind the Gamut API.

Host-side mirror of the reference interface for the hot path: PixelType, LoadFlags,
scanlinesConvert, Image.loadFromMemory / Image.convertTo. All pixel work runs in the CUDA
library gamut_b200/libgamut_b200.so (C ABI in include/gamut_b200.h). No CPU fallback.
"""
from .types import *  # noqa: F401,F403
from .types import PixelType, ImageFormat, pixelTypeSize  # noqa: F401
from ._lib import GamutB200Error, last_error  # noqa: F401
from .scanline import scanlinesConvert, scanlinesCopy, scanlinesInterType, scanlinesConvertDevice  # noqa: F401
