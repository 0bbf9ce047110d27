"""The QOI-10b encoder's kernels without a GPU: gamut_b200/csrc/qoi10b_encode.cuh compiled for the host under the
thread-per-CUDA-thread emulation (tests/cuda_emu.h, tests/emu_qoi10b_encode.cpp: the launches of
gb::qoi10b_encode_device) and compared, byte for byte, with the oracle's restatement of qoi10b_encode
(qoi10b.d:136-500); the streams also decode back to the image through the oracle's decoder."""
import ctypes as C
import os

import numpy as np
import pytest

from test_qoix_encode_emulated import Desc, expand

HERE = os.path.dirname(os.path.abspath(__file__))
SRCS = [os.path.join(HERE, "emu_qoi10b_encode.cpp"), os.path.join(HERE, "cuda_emu.h"),
        os.path.join(HERE, "..", "gamut_b200", "csrc", "qoi10b_encode.cuh"), os.path.join(HERE, "..", "gamut_b200", "csrc", "qoix_encode.cuh")]


@pytest.fixture(scope="module")
def emu():
    import emu_build
    return emu_build.build("emu_qoi10b_encode", SRCS)


def emu_encode(L, imgs, colorspace=0, par=-1.0, dpi=-1.0, descs=None):
    n = len(imgs)
    P, O, LN, D = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_int * n)(), (Desc * n)()
    keep, outs = [], []
    for i, im in enumerate(imgs):
        a = np.ascontiguousarray(im)
        keep.append(a)
        if descs is not None:
            D[i] = descs[i]
            h, w = D[i].height, D[i].width
        else:
            h, w, c = a.shape
            D[i] = Desc(w, h, w * c * 2, c, 10, colorspace, 0, par, dpi)
        P[i] = a.ctypes.data
        cap = (w * h * 52 + 7) // 8 + 25 + 5 + 64
        buf = np.full(cap + 32, 0xEE, np.uint8)
        o = (-buf.ctypes.data) % 16
        keep.append(buf)
        O[i] = buf.ctypes.data + o
        outs.append((buf, o, cap))
    assert L.emu_qoi10b_encode_batch(n, P, D, O, LN) == 1
    res = []
    for i, (buf, o, cap) in enumerate(outs):
        ln = LN[i]
        assert ln <= cap
        if ln > 0:
            assert (buf[o + ln + 3:] == 0xEE).all()
        res.append(bytes(buf[o:o + ln]) if ln > 0 else None)
    return res


def qoi10b_images(c, rng):
    h, w = 70, 91
    imgs = [expand(rng.integers(0, 1024, (hh, ww, c))) for (hh, ww) in [(1, 1), (1, 2), (2, 1), (3, 5), (33, 47), (2, 300)]]   # noise: RGB / RGBA
    imgs.append(expand(np.cumsum(rng.integers(-2, 3, (h, w, c)), axis=1) % 1024))         # LUMA0 / LUMA / ADIFF
    imgs.append(expand(np.cumsum(rng.integers(-20, 21, (h, w, c)), axis=1) % 1024))       # LUMA2 / ADIFF2
    imgs.append(expand(np.cumsum(rng.integers(-90, 91, (h, w, c)), axis=0) % 1024))       # LUMA3, vertical structure: the average predictor
    g = rng.integers(0, 1024, (h, w, 1))
    imgs.append(expand(np.concatenate([np.repeat(g, 3, axis=2), np.full((h, w, c - 3), 1023)], axis=2)))   # GRAY
    v = np.zeros((h * w, c), np.int64)                                                    # runs around 7 / 8 / 256 / tiles
    pos = 0
    for n in [1, 1, 2, 7, 8, 9, 255, 256, 257, 258, 512, 513, 1, 3, 2100]:
        if pos >= h * w:
            break
        v[pos:pos + n] = rng.integers(0, 1024, c)
        pos += n
    v[pos:] = rng.integers(0, 1024, (max(h * w - pos, 0), c))
    imgs.append(expand(v.reshape(h, w, c)))
    imgs.append(expand(np.full((40, 130, c), 517)))                                       # one flat image
    first = np.zeros((40, 130, c), np.int64)
    first[..., 3:] = 1023
    imgs.append(expand(first))                                                            # equal to the initial pixel
    imgs.append(expand(rng.integers(0, 3, (60, 60, c)) * 400))                            # few colours: short runs and big steps
    return imgs


@pytest.mark.parametrize("c", [3, 4])
def test_streams_equal_the_oracle(emu, oracle, c):
    rng = np.random.default_rng(50 + c)
    imgs = qoi10b_images(c, rng)
    got = emu_encode(emu, imgs, colorspace=1, par=1.5, dpi=96.0)
    for im, g in zip(imgs, got):
        exp = oracle.qoi10b_encode(im, colorspace=1, par=1.5, dpi=96.0)
        assert exp is not None and g == exp
        assert np.array_equal(oracle.qoix_decode(g, 0)[0], im)     # the reference's round-trip property


def test_pitch_and_rejects(emu, oracle):
    rng = np.random.default_rng(3)
    img = expand(rng.integers(0, 1024, (20, 30, 4)))
    wide = rng.integers(0, 65536, (20, 37, 4)).astype(np.uint16)
    wide[:, :30] = img                                             # row padding must not be read as pixels
    img3 = expand(np.cumsum(rng.integers(-9, 10, (20, 30, 3)), axis=1) % 1024)
    descs = [Desc(30, 20, 296, 4, 10, 0, 0, -1.0, -1.0), Desc(30, 20, 180, 3, 10, 2, 0, -1.0, -1.0), Desc(30, 20, 296, 4, 8, 0, 0, -1.0, -1.0),
             Desc(30, 20, 296, 2, 10, 0, 0, -1.0, -1.0), Desc(30, 20, 296, 4, 10, 0, 1, -1.0, -1.0), Desc(30, 20, 238, 4, 10, 0, 0, -1.0, -1.0),
             Desc(0, 20, 296, 4, 10, 0, 0, -1.0, -1.0), Desc(30, 20, 297, 4, 10, 0, 0, -1.0, -1.0)]
    got = emu_encode(emu, [wide, img3, wide, wide, wide, wide, wide, wide], descs=descs)
    assert got[0] == oracle.qoi10b_encode(img) and got[1] == oracle.qoi10b_encode(img3, colorspace=2)
    assert got[2:] == [None] * 6
