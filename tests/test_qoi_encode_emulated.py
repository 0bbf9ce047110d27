"""The QOI encoder's kernels without a GPU: gamut_b200/csrc/qoi_encode.cuh is compiled for the host under a
thread-per-CUDA-thread emulation (tests/cuda_emu.h, tests/emu_qoi_encode.cpp: the same five launches as
gb::qoi_encode_device) and its streams are compared, byte for byte, with the oracle's restatement of qoi_encode
(qoi.d:295-426). This checks the text of the kernels -- indexing, scans, the per-bucket last-writer query, the byte
placement -- not the hardware; tests/test_qoi_encode_gpu.py is the parity test of the product path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from qoixutil import qoi_test_image

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libemu_qoi_encode.so")
SRCS = [os.path.join(HERE, "emu_qoi_encode.cpp"), os.path.join(HERE, "cuda_emu.h"),
        os.path.join(HERE, "..", "gamut_b200", "csrc", "qoi_encode.cuh")]


@pytest.fixture(scope="module")
def emu():
    import emu_build
    L = emu_build.build("emu_qoi_encode", SRCS)
    return L


def emu_encode(L, imgs, pitches=None, colorspace=0, misalign=0):
    """imgs: (h, w, c) uint8 arrays, or (array, first-scanline byte offset, h, w, c) for explicit layouts."""
    n = len(imgs)
    P, W, H = (C.c_void_p * n)(), (C.c_uint32 * n)(), (C.c_uint32 * n)()
    CH, CS, PI, O, LN = (C.c_int * n)(), (C.c_int * n)(), (C.c_int * n)(), (C.c_void_p * n)(), (C.c_int * n)()
    keep, outs = [], []
    for i, im in enumerate(imgs):
        if isinstance(im, tuple):
            a, off, h, w, c = im
        else:
            h, w, c = im.shape
            a, off = np.ascontiguousarray(im), 0
            if misalign:
                b = np.zeros(a.size + 8, np.uint8)
                o = (-b.ctypes.data) % 4 + misalign
                b[o:o + a.size] = a.reshape(-1)
                a, off = b, o
        keep.append(a)
        P[i], W[i], H[i], CH[i], CS[i] = a.ctypes.data + off, w, h, c, colorspace
        PI[i] = w * c if pitches is None else pitches[i]
        cap = w * h * (c + 1) + 14 + 8 + 16
        buf = np.full(cap + 32, 0xEE, np.uint8)
        o = (-buf.ctypes.data) % 16
        keep.append(buf)
        O[i] = buf.ctypes.data + o
        outs.append((buf, o, cap))
    assert L.emu_qoi_encode_batch(n, P, W, H, CH, CS, PI, O, LN) == 1
    res = []
    for i, (buf, o, cap) in enumerate(outs):
        ln = LN[i]
        assert ln <= cap
        if ln > 0:
            assert (buf[o + ln + 3:] == 0xEE).all()            # nothing written past the stream (word stores: + 3)
        res.append(bytes(buf[o:o + ln]) if ln > 0 else None)
    return res


def test_streams_equal_the_oracle(emu, oracle):
    rng = np.random.default_rng(5)
    imgs = [qoi_test_image(40, 50, 4, 1), qoi_test_image(33, 70, 3, 2), np.zeros((6, 40, 4), np.uint8),
            np.zeros((3, 70, 3), np.uint8), rng.integers(0, 4, (30, 30, 4)).astype(np.uint8) * 60,
            np.zeros((1, 1, 4), np.uint8), np.full((1, 1, 3), 7, np.uint8), qoi_test_image(70, 91, 4, 3),
            rng.integers(0, 256, (50, 45, 4)).astype(np.uint8), rng.integers(0, 2, (64, 80, 3)).astype(np.uint8) * 255]
    s = np.zeros((4, 30, 4), np.uint8); s[..., 3] = 255; s[2:, :, 0] = 9          # starts inside the initial run
    imgs.append(s)
    v = np.zeros((70 * 91, 4), np.uint8)                                          # runs around 62 and across tiles
    pos = 0
    for n in [1, 1, 2, 61, 62, 63, 124, 125, 300, 1, 3, 1024, 1100]:
        v[pos:pos + n] = rng.integers(0, 256, 4)
        pos += n
    v[pos:] = rng.integers(0, 256, (70 * 91 - pos, 4))
    imgs.append(v.reshape(70, 91, 4))
    pal = rng.integers(0, 256, (5, 4)).astype(np.uint8)                           # index hits within and across tiles
    imgs.append(pal[rng.integers(0, 5, 60 * 60)].reshape(60, 60, 4))
    pal3 = rng.integers(0, 256, (90, 3)).astype(np.uint8)                         # buckets shared by several colours
    imgs.append(pal3[rng.integers(0, 90, 60 * 60)].reshape(60, 60, 3))
    z = rng.integers(0, 256, (40, 40, 4)).astype(np.uint8); z[5::7] = 0           # the all-zero pixel against the zeroed index
    imgs.append(z)
    got = emu_encode(emu, imgs, colorspace=1)
    for im, g in zip(imgs, got):
        assert g == oracle.qoi_encode(im, colorspace=1)


def test_pitch_alignment_and_rejects(emu, oracle):
    rng = np.random.default_rng(9)
    img = qoi_test_image(37, 45, 4, 4)
    exp = oracle.qoi_encode(img)
    # padded rows: the padding must not be read as pixels
    wide = rng.integers(0, 256, (37, 60, 4)).astype(np.uint8)
    wide[:, :45] = img
    # vertically flipped storage: first scanline = last row of the buffer, negative pitch (saveQOI passes image._pitch)
    flipped = np.ascontiguousarray(wide[::-1])
    got = emu_encode(emu, [(wide, 0, 37, 45, 4), (flipped, 36 * 240, 37, 45, 4)], pitches=[240, -240])
    assert got[0] == exp and got[1] == exp
    # rgba8 at an address that is not a multiple of 4 (byte loads), rgb8 with an odd pitch
    assert emu_encode(emu, [img], misalign=1)[0] == exp
    img3 = qoi_test_image(20, 31, 3, 6)
    wide3 = rng.integers(0, 256, (20, 31 * 3 + 5), dtype=np.uint8)
    wide3[:, :93] = img3.reshape(20, 93)
    assert emu_encode(emu, [(wide3, 0, 20, 31, 3)], pitches=[98])[0] == oracle.qoi_encode(img3)
    # qoi_encode's own refusals (qoi.d:303-311) and a pitch smaller than a scanline
    ok = qoi_test_image(4, 4, 4, 1)
    bad = emu_encode(emu, [(ok, 0, 4, 4, 2), (ok, 0, 4, 4, 5), (ok, 0, 0, 4, 4), (ok, 0, 4, 4, 4), ok], pitches=[8, 20, 16, 8, 16])
    assert bad[:4] == [None] * 4 and bad[4] == oracle.qoi_encode(ok)
    assert emu_encode(emu, [ok], colorspace=2) == [None]
