"""The BMP writer without a GPU: gamut_b200/csrc/bmp_encode.cuh (header + kernel) compiled for the host under the
thread-per-CUDA-thread emulation and compared, byte for byte, with the oracle's restatement of write_bmp
(codecs/bmpenc.d:25-113); the oracle's files are read back by PIL's independent BMP reader and by the oracle's own."""
import ctypes as C
import io
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRCS = [os.path.join(HERE, "emu_bmp_encode.cpp"), os.path.join(HERE, "cuda_emu.h"), os.path.join(HERE, "..", "gamut_b200", "csrc", "bmp_encode.cuh")]


@pytest.fixture(scope="module")
def emu():
    import emu_build
    L = emu_build.build("emu_bmp_encode", SRCS)
    L.emu_bmp_encode.restype = C.c_long
    L.emu_bmp_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_size_t]
    return L


def emu_encode(L, img, ppmX=-1.0, ppmY=-1.0, pitch=None, first_scanline=0, shape=None, type_=None):
    px = np.ascontiguousarray(img)
    h, w, c = shape if shape is not None else px.shape
    t = type_ if type_ is not None else {3: 9, 4: 12}.get(c, -1)
    out = np.full(122 + h * (w * c + 3) + 64, 0xEE, np.uint8)
    n = L.emu_bmp_encode(px.ctypes.data + first_scanline, t, w, h, pitch if pitch is not None else w * c, ppmX, ppmY, out.ctypes.data, out.size)
    assert n >= 0
    if n == 0:
        return None
    assert (out[n:] == 0xEE).all()
    return out[:n].tobytes()


@pytest.mark.parametrize("c", [3, 4])
def test_files_equal_the_oracle_and_read_back(emu, oracle, c):
    from PIL import Image as PILImage
    rng = np.random.default_rng(c)
    for (h, w) in [(1, 1), (1, 2), (2, 3), (5, 121), (7, 122), (9, 123), (3, 257), (4, 600), (33, 47)]:       # every padding, headers narrower / wider than a row
        img = rng.integers(0, 256, (h, w, c)).astype(np.uint8)
        exp = oracle.bmp_encode(img, ppmX=3779.53, ppmY=7874.4)
        assert exp is not None and emu_encode(emu, img, ppmX=3779.53, ppmY=7874.4) == exp
        assert np.array_equal(np.asarray(PILImage.open(io.BytesIO(exp)).convert("RGBA" if c == 4 else "RGB")), img)
        back = oracle.bmp_load(exp, 0)
        assert back is not None and np.array_equal(back[0], img) and abs(back[2] - 3780) < 1e-3 and abs(back[3] - 7874) < 1e-3
    assert oracle.bmp_encode(img)[38:46] == b"\0" * 8               # unknown resolution: 0 pixels per metre


def test_pitch_flip_and_rejects(emu, oracle):
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (21, 45, 3)).astype(np.uint8)
    exp = oracle.bmp_encode(img)
    wide = rng.integers(0, 256, (21, 60, 3)).astype(np.uint8)
    wide[:, :45] = img
    assert emu_encode(emu, wide, pitch=180, shape=(21, 45, 3)) == exp
    flipped = np.ascontiguousarray(wide[::-1])
    assert emu_encode(emu, flipped, pitch=-180, first_scanline=20 * 180, shape=(21, 45, 3)) == exp
    assert oracle.bmp_encode(flipped, pitch=-180, first_scanline=20 * 180, shape=(21, 45, 3)) == exp
    for kw in ({"type_": 0}, {"type_": 13}, {"shape": (21, 0, 3)}, {"shape": (0, 45, 3)}, {"shape": (1, 32768, 3)}):
        assert emu_encode(emu, img, **kw) is None and oracle.bmp_encode(img, **kw) is None
    assert emu_encode(emu, img, pitch=100) is None
