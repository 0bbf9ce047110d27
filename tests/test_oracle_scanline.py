"""Oracle (oracle/scanline_oracle.c) vs independent numpy float32 closed forms of scanline.d.

The reference has no test that pins a converted value (parity unpinned); these
checks validate the restatement against the formulas read from the source.
"""
import numpy as np
import pytest

from gamut_b200.types import PixelType as PT, pixelTypeSize

f32 = np.float32


def conv(oracle, s, d, src, w, h=1):
    dst = np.zeros(w * h * pixelTypeSize(d), np.uint8)
    assert oracle.scanlines_convert(s, src.view(np.uint8).reshape(-1), w * pixelTypeSize(s), d, dst,
                                    w * pixelTypeSize(d), w, h)
    return dst


def test_sizes_and_intertype(oracle):
    L = oracle.lib()
    for t in range(18):
        assert L.or_pixelTypeSize(t) == pixelTypeSize(t)
    eight = {PT.l8, PT.la8, PT.rgb8, PT.rgba8}
    for s in range(18):
        for d in range(18):
            exp = PT.rgba8 if (s in eight and d in eight) else PT.rgbaf32
            assert L.or_scanlinesInterType(s, d) == exp


def test_rgba8_to_rgbaf32_all_values(oracle):
    src = np.arange(256, dtype=np.uint8).repeat(4)
    out = conv(oracle, PT.rgba8, PT.rgbaf32, src, 256).view(f32)
    exp = (np.arange(256).astype(f32) / f32(255.0)).repeat(4)
    assert np.array_equal(out.view(np.uint32), exp.view(np.uint32))


def test_rgbaf32_to_rgba8_roundtrip_and_formula(oracle):
    rng = np.random.default_rng(2)
    v = rng.random(4 * 4096, dtype=f32)
    out = conv(oracle, PT.rgbaf32, PT.rgba8, v, 4096)
    exp = (f32(0.5) + v * f32(255.0)).astype(np.int32).astype(np.uint8)  # trunc toward zero
    assert np.array_equal(out, exp)
    # u8 -> f32 -> u8 is the identity
    src = np.arange(256, dtype=np.uint8).repeat(4)
    f = conv(oracle, PT.rgba8, PT.rgbaf32, src, 256)
    back = conv(oracle, PT.rgbaf32, PT.rgba8, f, 256)
    assert np.array_equal(back, src)


def test_grey_formula_left_assoc(oracle):
    rng = np.random.default_rng(3)
    v = rng.random(4 * 1000, dtype=f32).reshape(-1, 4)
    out = conv(oracle, PT.rgbaf32, PT.l16, v, 1000).view(np.uint16)
    s = (v[:, 0] + v[:, 1]) + v[:, 2]
    exp = (f32(0.5) + (s * f32(65535.0)) / f32(3.0)).astype(np.int32).astype(np.uint16)
    assert np.array_equal(out, exp)
    outp = conv(oracle, PT.rgbaf32, PT.lap8, v, 1000).reshape(-1, 2)
    expl = (f32(0.5) + ((s * v[:, 3]) * f32(255.0)) / f32(3.0)).astype(np.int32).astype(np.uint8)
    expa = (f32(0.5) + v[:, 3] * f32(255.0)).astype(np.int32).astype(np.uint8)
    assert np.array_equal(outp[:, 0], expl) and np.array_equal(outp[:, 1], expa)


def test_unpremultiply_guard(oracle):
    src = np.array([10, 20, 30, 0, 10, 20, 30, 128], np.uint8)
    out = conv(oracle, PT.rgbap8, PT.rgbaf32, src, 2).view(f32).reshape(2, 4)
    c = np.array([10, 20, 30], f32) / f32(255)
    a = f32(128) / f32(255)
    assert np.array_equal(out[0], np.array([c[0], c[1], c[2], 0], f32))
    assert np.array_equal(out[1], np.array([c[0] / a, c[1] / a, c[2] / a, a], f32))


def test_8bit_block_is_integer_and_takes_red(oracle):
    src = np.array([1, 2, 3, 4, 5, 6, 7, 8], np.uint8)
    assert conv(oracle, PT.rgba8, PT.l8, src, 2).tolist() == [1, 5]          # R channel, scanline.d:201
    assert conv(oracle, PT.rgba8, PT.la8, src, 2).tolist() == [1, 4, 5, 8]
    assert conv(oracle, PT.rgba8, PT.rgb8, src, 2).tolist() == [1, 2, 3, 5, 6, 7]
    assert conv(oracle, PT.l8, PT.rgba8, src[:2], 2).tolist() == [1, 1, 1, 255, 2, 2, 2, 255]
    assert conv(oracle, PT.la8, PT.rgb8, src[:4], 2).tolist() == [1, 1, 1, 3, 3, 3]


def test_int_to_int_through_float(oracle):
    # l16 -> l8 goes through rgbaf32: (u8)(0.5 + ((3*(v/65535))*255)/3)
    v = np.arange(0, 65536, 7, dtype=np.uint16)
    out = conv(oracle, PT.l16, PT.l8, v, len(v))
    b = v.astype(f32) / f32(65535.0)
    s = (b + b) + b
    exp = (f32(0.5) + (s * f32(255.0)) / f32(3.0)).astype(np.int32).astype(np.uint8)
    assert np.array_equal(out, exp)


def test_negative_and_padded_pitch(oracle):
    w, h = 5, 4
    rng = np.random.default_rng(5)
    src = rng.integers(0, 256, (h, w * 4 + 3), dtype=np.uint8)
    dst = np.full((h, w * 16 + 8), 0xAB, np.uint8)
    # read bottom-up, write top-down
    assert oracle.scanlines_convert(PT.rgba8, src.reshape(-1), -(w * 4 + 3), PT.rgbaf32, dst.reshape(-1),
                                    w * 16 + 8, w, h, src_off=(h - 1) * (w * 4 + 3))
    for y in range(h):
        exp = src[h - 1 - y, :w * 4].astype(f32) / f32(255)
        assert np.array_equal(dst[y, :w * 16].view(f32), exp)
        assert (dst[y, w * 16:] == 0xAB).all()
