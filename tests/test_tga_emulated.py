"""The TGA decoder's header walk and kernels without a GPU: gamut_b200/csrc/tga.cuh compiled for the host under the
thread-per-CUDA-thread emulation (tests/cuda_emu.h, tests/emu_tga.cpp) and compared with the oracle's restatement of
TGADecoder (codecs/tga.d:313-588) on every variant of tests/tgautil.py, truncated files included."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tgautil import make_tga, pil_tga
from test_oracle_tga import CASES

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libemu_tga.so")
SRCS = [os.path.join(HERE, "emu_tga.cpp"), os.path.join(HERE, "cuda_emu.h"), os.path.join(HERE, "..", "gamut_b200", "csrc", "tga.cuh")]


@pytest.fixture(scope="module")
def emu():
    import emu_build
    L = emu_build.build("emu_tga", SRCS)
    L.emu_tga_load.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t] + [C.POINTER(C.c_int)] * 3
    return L


def emu_load(L, data):
    out = np.full(1 << 20, 0xEE, np.uint8)
    w, h, c = C.c_int(), C.c_int(), C.c_int()
    r = L.emu_tga_load(data, len(data), out.ctypes.data, out.size, C.byref(w), C.byref(h), C.byref(c))
    assert r >= 0
    if r == 0:
        return None
    n = w.value * h.value * c.value
    assert (out[n:] == 0xEE).all()
    return out[:n].reshape(h.value, w.value, c.value).copy()


def same(a, b):
    return (a is None and b is None) or (a is not None and b is not None and a.shape == b.shape and np.array_equal(a, b))


@pytest.mark.parametrize("kind,kw", CASES)
def test_variants(emu, oracle, kind, kw):
    rng = np.random.default_rng(7)
    for rle in (False, True):
        for top_down in (False, True):
            for (w, h) in [(1, 1), (13, 9), (131, 40)]:
                data, _ = make_tga(w, h, kind, rng, rle=rle, top_down=top_down, **kw)
                exp = oracle.tga_load(data)
                assert exp is not None and same(emu_load(emu, data), exp)
                for cut in (len(data) - 1, len(data) - 7, len(data) // 2, 19, 18, 5):
                    assert same(emu_load(emu, data[:cut]), oracle.tga_load(data[:cut]))


def test_pil_files_overrun_and_rejects(emu, oracle):
    rng = np.random.default_rng(1)
    for c in (1, 3, 4):
        img = rng.integers(0, 4, (23, 37, c)).astype(np.uint8) * 80
        for rle in (False, True):
            data = pil_tga(img, rle, bool(c & 1))
            assert same(emu_load(emu, data), img)
    data, _ = make_tga(17, 11, "bgr24", rng, rle=True, overrun=True)
    assert same(emu_load(emu, data), oracle.tga_load(data))
    ok, _ = make_tga(5, 4, "pal24", rng, pal_len=9)
    for at, v in [(1, 2), (2, 4), (16, 12), (16, 24), (7, 12)]:
        bad = bytearray(ok); bad[at] = v
        assert emu_load(emu, bytes(bad)) is None and oracle.tga_load(bytes(bad)) is None
    for _ in range(300):                                        # header fuzz: both sides accept / reject the same files
        bad = bytearray(ok)
        for _ in range(int(rng.integers(1, 4))):
            bad[int(rng.integers(0, 18))] = int(rng.integers(0, 256))
        bad[12:16] = (5).to_bytes(2, "little") + (4).to_bytes(2, "little")      # keep the size small
        assert same(emu_load(emu, bytes(bad)), oracle.tga_load(bytes(bad)))


# ---- encoder (tga_encode.cuh) ------------------------------------------------------------------------------------------
def emu_encode(L, img, pitch=None, first_scanline=0, shape=None):
    px = np.ascontiguousarray(img)
    h, w, c = shape if shape is not None else px.shape
    t = {1: 0, 2: 3, 3: 9, 4: 12}[c]
    out = np.full(18 + h * (w * 5 + 2) + 64, 0xEE, np.uint8)
    L.emu_tga_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    n = L.emu_tga_encode(px.ctypes.data + first_scanline, t, w, h, pitch if pitch is not None else w * c, out.ctypes.data, out.size)
    assert n >= 0
    if n == 0:
        return None
    assert (out[n:] == 0xEE).all()
    return out[:n].tobytes()


def tga_encode_images(c, rng):
    """Images that exercise the packet choice: noise (raw packets of 128), flat (runs of 128), runs and raw stretches of
    every length around 1 / 2 / 127 / 128 / 129, rows that end inside a run or a raw stretch, width 1."""
    imgs = [rng.integers(0, 256, (5, 300, c)).astype(np.uint8), np.full((4, 300, c), 9, np.uint8), np.zeros((3, 1, c), np.uint8)]
    imgs.append((rng.integers(0, 3, (17, 300, c)) * 100).astype(np.uint8))
    rows = []
    for seed in range(24):
        row = []
        while len(row) < 400:
            n = int(rng.choice([1, 1, 2, 3, 126, 127, 128, 129, 130, 255, 256, 257, 5]))
            if rng.random() < 0.5:
                row += [rng.integers(0, 256, c)] * n                                # a run
            else:
                row += list(rng.integers(0, 256, (n, c)))                           # a stretch of (mostly) different pixels
        rows.append(np.array(row[:400 - seed % 3], np.uint8))
    for k in range(3):
        imgs.append(np.stack([r[:397] for r in rows[k::3]]))
    return imgs


@pytest.mark.parametrize("c", [1, 2, 3, 4])
def test_encoder_files_equal_the_oracle(emu, oracle, c):
    rng = np.random.default_rng(20 + c)
    for img in tga_encode_images(c, rng):
        exp = oracle.tga_encode(img)
        got = emu_encode(emu, img)
        assert exp is not None and got == exp
        dec = oracle.tga_load(got)                                # and the file decodes back to the image (as rgb8 / rgba8)
        rgb = img if c >= 3 else np.concatenate([np.repeat(img[..., :1], 3, axis=2), img[..., 1:]], axis=2)
        assert np.array_equal(dec, rgb)


def test_encoder_pitch_and_flip(emu, oracle):
    rng = np.random.default_rng(4)
    img = (rng.integers(0, 3, (21, 45, 4)) * 90).astype(np.uint8)
    exp = oracle.tga_encode(img)
    wide = rng.integers(0, 256, (21, 60, 4)).astype(np.uint8)
    wide[:, :45] = img
    assert emu_encode(emu, wide, pitch=240, shape=(21, 45, 4)) == exp
    flipped = np.ascontiguousarray(wide[::-1])
    assert emu_encode(emu, flipped, pitch=-240, first_scanline=20 * 240, shape=(21, 45, 4)) == exp
    assert oracle.tga_encode(flipped, pitch=-240, first_scanline=20 * 240, shape=(21, 45, 4)) == exp


def test_rle_stream_longer_than_the_staging_window(emu, oracle):
    """Run-length files several times the 8 KB window of tga_rle_kernel, at every byte alignment of the file start."""
    rng = np.random.default_rng(12)
    for kind, kw in [("bgra32", {}), ("bgr24", {"idlen": 5}), ("pal16", {"pal_len": 300, "index_bits": 16}), ("l8", {})]:
        data, _ = make_tga(300, 90, kind, rng, rle=True, **kw)
        assert len(data) > 20000 or kind in ("l8",)
        exp = oracle.tga_load(data)
        for shift in range(4):
            buf = np.zeros(len(data) + 8, np.uint8)
            o = (-buf.ctypes.data) % 4 + shift
            buf[o:o + len(data)] = np.frombuffer(data, np.uint8)
            out = np.full(1 << 20, 0xEE, np.uint8)
            w, h, c = C.c_int(), C.c_int(), C.c_int()
            emu.emu_tga_load.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t] + [C.POINTER(C.c_int)] * 3
            assert emu.emu_tga_load(buf.ctypes.data + o, len(data), out.ctypes.data, out.size, C.byref(w), C.byref(h), C.byref(c)) == 1
            assert np.array_equal(out[: exp.size].reshape(exp.shape), exp)
        emu.emu_tga_load.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t] + [C.POINTER(C.c_int)] * 3
        for cut in (len(data) - 1, len(data) - 600, 8192 + 18, 9000):
            assert same(emu_load(emu, data[:cut]), oracle.tga_load(data[:cut]))
