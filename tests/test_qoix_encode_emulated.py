"""The QOI-Plane10 / QOI-Plane encoder kernels without a GPU: gamut_b200/csrc/qoix_encode.cuh compiled for the host
under the thread-per-CUDA-thread emulation (tests/cuda_emu.h, tests/emu_qoix_encode.cpp: the launches of
gb::qoiplane_encode_device) and compared, byte for byte, with the oracle's restatements of qoiplane10_encode
(qoiplane10.d:99-314) and qoiplane_encode (qoiplane.d:109-375). The QOI-Plane10 kernels were verified on a B200 before
this test existed, so the 10-bit half also checks the emulation itself."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from qoixutil import depth_map_la

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libemu_qoix_encode.so")
SRCS = [os.path.join(HERE, "emu_qoix_encode.cpp"), os.path.join(HERE, "cuda_emu.h"),
        os.path.join(HERE, "..", "gamut_b200", "csrc", "qoix_encode.cuh")]


class Desc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("pitchBytes", C.c_int32), ("channels", C.c_uint8),
                ("bitdepth", C.c_uint8), ("colorspace", C.c_uint8), ("compression", C.c_uint8),
                ("pixelAspectRatio", C.c_float), ("resolutionY", C.c_float)]


@pytest.fixture(scope="module")
def emu():
    import emu_build
    L = emu_build.build("emu_qoix_encode", SRCS)
    return L


def emu_encode(L, imgs, colorspace=0, par=-1.0, dpi=-1.0, pitches=None, descs=None):
    """imgs: (h, w, c) uint8 (QOI-Plane) or uint16 (QOI-Plane10) arrays; one batch, mixed."""
    n = len(imgs)
    P, O, LN, D = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_int * n)(), (Desc * n)()
    keep, outs = [], []
    for i, im in enumerate(imgs):
        a = np.ascontiguousarray(im)
        keep.append(a)
        if descs is not None:
            D[i] = descs[i]
            h, w, c = D[i].height, D[i].width, max(int(D[i].channels), 1)
        else:
            h, w, c = a.shape
            D[i] = Desc(w, h, pitches[i] if pitches else w * c * a.itemsize, c, 10 if a.itemsize == 2 else 8, colorspace, 0, par, dpi)
        P[i] = a.ctypes.data
        cap = (w * h * 28 + 7) // 8 + 25 + 5 + 64
        buf = np.full(cap + 32, 0xEE, np.uint8)
        o = (-buf.ctypes.data) % 16
        keep.append(buf)
        O[i] = buf.ctypes.data + o
        outs.append((buf, o))
    assert L.emu_qoix_encode_batch(n, P, D, O, LN) == 1
    return [bytes(buf[o:o + LN[i]]) if LN[i] > 0 else None for i, (buf, o) in enumerate(outs)]


def expand(v):
    return ((v << 6) | (v >> 4)).astype(np.uint16)


def plane8_images(c, rng):
    h, w = 70, 91
    imgs = [(depth_map_la(hh, ww, 3 + hh, c) >> 8).astype(np.uint8) for (hh, ww) in [(1, 1), (1, 2), (2, 1), (3, 5), (33, 47), (2, 300), (64, 64)]]
    imgs.append(rng.integers(0, 256, (h, w, c)).astype(np.uint8))                                 # noise: DIRECT / LA
    imgs.append((np.cumsum(rng.integers(-6, 7, (h, w, c)), axis=1) % 256).astype(np.uint8))      # DIFF1 / DIFF2 / ADIFF
    imgs.append((np.cumsum(rng.integers(-2, 3, (h, w, c)), axis=0) % 256).astype(np.uint8))      # vertical structure: the average predictor
    v = np.zeros((h * w, c), np.uint8)                                                            # runs around 3 / 4 / 258 / tiles
    pos = 0
    for n in [1, 1, 2, 3, 4, 5, 257, 258, 259, 260, 516, 517, 1, 3, 1100]:
        v[pos:pos + n] = rng.integers(0, 256, c)
        pos += n
    v[pos:] = rng.integers(0, 256, (h * w - pos, c))
    imgs.append(v.reshape(h, w, c))
    imgs.append(np.full((40, 130, c), 117, np.uint8))                                             # one flat image
    first = np.zeros((40, 130, c), np.uint8)
    first[..., 1:] = 255
    imgs.append(first)                                                                            # equal to the initial predictor
    return imgs


@pytest.mark.parametrize("c", [1, 2])
def test_qoiplane_streams_equal_the_oracle(emu, oracle, c):
    rng = np.random.default_rng(10 + c)
    imgs = plane8_images(c, rng)
    got = emu_encode(emu, imgs, colorspace=1, par=1.5, dpi=96.0)
    for im, g in zip(imgs, got):
        exp = oracle.qoiplane_encode(im, colorspace=1, par=1.5, dpi=96.0)
        assert exp is not None and g == exp
        dec = oracle.qoix_decode(g, 0)                              # the reference's round-trip property
        assert np.array_equal(dec[0], im)


@pytest.mark.parametrize("c", [1, 2])
def test_qoiplane10_streams_equal_the_oracle(emu, oracle, c):
    rng = np.random.default_rng(c)
    imgs = [depth_map_la(h, w, 3 + h, c) for (h, w) in [(1, 1), (3, 5), (33, 47), (2, 300), (64, 64)]]
    imgs.append(expand(rng.integers(0, 1024, (70, 91, c))))
    imgs.append(expand(np.cumsum(rng.integers(-40, 41, (70, 91, c)), axis=1) % 1024))
    imgs.append(expand(np.full((40, 130, c), 517)))
    got = emu_encode(emu, imgs, par=1.5, dpi=96.0)
    for im, g in zip(imgs, got):
        assert g == oracle.qoiplane10_encode(im, par=1.5, dpi=96.0)


def test_mixed_batch_pitch_and_rejects(emu, oracle):
    rng = np.random.default_rng(3)
    a8 = (depth_map_la(20, 30, 1, 2) >> 8).astype(np.uint8)
    a10 = depth_map_la(20, 30, 2, 2)
    wide8 = rng.integers(0, 256, (20, 37, 2)).astype(np.uint8)
    wide8[:, :30] = a8                                              # row padding must not be read as pixels
    descs = [Desc(30, 20, 74, 2, 8, 0, 0, -1.0, -1.0), Desc(30, 20, 120, 2, 10, 0, 0, -1.0, -1.0),
             Desc(30, 20, 74, 3, 8, 0, 0, -1.0, -1.0), Desc(30, 20, 74, 2, 8, 0, 1, -1.0, -1.0),
             Desc(30, 20, 59, 2, 8, 0, 0, -1.0, -1.0), Desc(0, 20, 74, 2, 8, 0, 0, -1.0, -1.0),
             Desc(30, 20, 60, 2, 8, 0, 0, -1.0, -1.0)]
    got = emu_encode(emu, [wide8, a10, wide8, wide8, wide8, wide8, a8], descs=descs)
    assert got[0] == oracle.qoiplane_encode(a8) and got[1] == oracle.qoiplane10_encode(a10) and got[6] == got[0]
    assert got[2:6] == [None] * 4
