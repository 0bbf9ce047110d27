"""The QOI2AVG encoder's kernels without a GPU: gamut_b200/csrc/qoi2avg_encode.cuh compiled for the host under the
thread-per-CUDA-thread emulation (tests/cuda_emu.h, tests/emu_qoi2avg_encode.cpp: the launches of
gb::qoi2avg_encode_device) and compared, byte for byte, with the oracle's restatement of qoix_encode
(qoi2avg.d:376-617); the streams also decode back to the image through the oracle's decoder."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from qoixutil import qoi_test_image
from test_qoix_encode_emulated import Desc

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libemu_qoi2avg_encode.so")
SRCS = [os.path.join(HERE, "emu_qoi2avg_encode.cpp"), os.path.join(HERE, "cuda_emu.h"),
        os.path.join(HERE, "..", "gamut_b200", "csrc", "qoi2avg_encode.cuh"), os.path.join(HERE, "..", "gamut_b200", "csrc", "qoi_encode.cuh")]


@pytest.fixture(scope="module")
def emu():
    import emu_build
    L = emu_build.build("emu_qoi2avg_encode", SRCS)
    return L


def emu_encode(L, imgs, colorspace=0, par=-1.0, dpi=-1.0, descs=None):
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
            D[i] = Desc(w, h, w * c, c, 8, colorspace, 0, par, dpi)
        P[i] = a.ctypes.data
        cap = w * h * (c + 1) + 25 + 4 + 16
        buf = np.full(cap + 32, 0xEE, np.uint8)
        o = (-buf.ctypes.data) % 16
        keep.append(buf)
        O[i] = buf.ctypes.data + o
        outs.append((buf, o, cap))
    assert L.emu_qoi2avg_encode_batch(n, P, D, O, LN) == 1
    res = []
    for i, (buf, o, cap) in enumerate(outs):
        ln = LN[i]
        assert ln <= cap
        if ln > 0:
            assert (buf[o + ln + 3:] == 0xEE).all()
        res.append(bytes(buf[o:o + ln]) if ln > 0 else None)
    return res


def qoi2avg_images(c, rng):
    h, w = 70, 91
    imgs = [qoi_test_image(hh, ww, c, 3 + hh) for (hh, ww) in [(1, 1), (1, 2), (2, 1), (3, 5), (33, 47), (2, 300), (64, 64)]]
    imgs.append(rng.integers(0, 256, (h, w, c)).astype(np.uint8))                                     # noise: RGB / RGBA
    imgs.append((np.cumsum(rng.integers(-2, 3, (h, w, c)), axis=1) % 256).astype(np.uint8))          # LUMA / ADIFF
    imgs.append((np.cumsum(rng.integers(-9, 10, (h, w, c)), axis=1) % 256).astype(np.uint8))         # LUMA2
    imgs.append((np.cumsum(rng.integers(-40, 41, (h, w, c)), axis=0) % 256).astype(np.uint8))        # LUMA3, vertical structure: LOCO-I
    g = rng.integers(0, 256, (h, w, 1)).astype(np.uint8)
    imgs.append(np.concatenate([np.repeat(g, 3, axis=2), np.full((h, w, c - 3), 255, np.uint8)], axis=2))  # GRAY
    v = np.zeros((h * w, c), np.uint8)                                                                # runs around 8 / 9 / 1024 / tiles
    pos = 0
    for n in [1, 1, 2, 8, 9, 10, 1023, 1024, 1025, 1, 3, 2100]:
        if pos >= h * w:
            break
        v[pos:pos + n] = rng.integers(0, 256, c)
        pos += n
    v[pos:] = rng.integers(0, 256, (max(h * w - pos, 0), c))
    imgs.append(v.reshape(h, w, c))
    imgs.append(np.full((40, 130, c), 200, np.uint8))                                                 # one flat image
    first = np.zeros((40, 130, c), np.uint8)
    first[..., 3:] = 255
    imgs.append(first)                                                                                # equal to the initial pixel
    imgs.append(np.zeros((40, 130, c), np.uint8))                                                     # rgba8: hits the zeroed index
    for ncol in (5, 40, 64, 65, 90, 300):                                                             # the FIFO: below, at and above its 64 entries
        pal = rng.integers(0, 256, (ncol, c)).astype(np.uint8)
        imgs.append(pal[rng.integers(0, ncol, 90 * 90)].reshape(90, 90, c))
    cyc = rng.integers(0, 256, (66, c)).astype(np.uint8)                                              # a cycle of 66 colours: every entry dies just before it returns
    imgs.append(cyc[np.arange(80 * 80) % 66].reshape(80, 80, c))
    cyc = rng.integers(0, 256, (64, c)).astype(np.uint8)                                              # a cycle of 64: every entry survives
    imgs.append(cyc[np.arange(80 * 80) % 64].reshape(80, 80, c))
    return imgs


@pytest.mark.parametrize("c", [3, 4])
def test_streams_equal_the_oracle(emu, oracle, c):
    rng = np.random.default_rng(30 + c)
    imgs = qoi2avg_images(c, rng)
    got = emu_encode(emu, imgs, colorspace=1, par=1.5, dpi=96.0)
    for im, g in zip(imgs, got):
        exp = oracle.qoi2avg_encode(im, colorspace=1, par=1.5, dpi=96.0)
        assert exp is not None and g == exp
        assert np.array_equal(oracle.qoix_decode(g, 0)[0], im)     # the reference's round-trip property


def test_pitch_and_rejects(emu, oracle):
    rng = np.random.default_rng(3)
    img = qoi_test_image(20, 30, 4, 1)
    wide = rng.integers(0, 256, (20, 37, 4)).astype(np.uint8)
    wide[:, :30] = img                                             # row padding must not be read as pixels
    img3 = qoi_test_image(20, 30, 3, 2)
    descs = [Desc(30, 20, 148, 4, 8, 0, 0, -1.0, -1.0), Desc(30, 20, 90, 3, 8, 2, 0, -1.0, -1.0), Desc(30, 20, 148, 4, 10, 0, 0, -1.0, -1.0),
             Desc(30, 20, 148, 4, 8, 3, 0, -1.0, -1.0), Desc(30, 20, 148, 4, 8, 0, 1, -1.0, -1.0), Desc(30, 20, 119, 4, 8, 0, 0, -1.0, -1.0),
             Desc(0, 20, 148, 4, 8, 0, 0, -1.0, -1.0), Desc(30, 20, 148, 5, 8, 0, 0, -1.0, -1.0)]
    got = emu_encode(emu, [wide, img3, wide, wide, wide, wide, wide, wide], descs=descs)
    assert got[0] == oracle.qoi2avg_encode(img) and got[1] == oracle.qoi2avg_encode(img3, colorspace=2)
    assert got[2:] == [None] * 6
