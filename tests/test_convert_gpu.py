"""GPU parity: CUDA PixelType converters (through the C ABI) vs the CPU oracle, bit-exact.
Reference: source/gamut/scanline.d:70-930."""
import numpy as np
import pytest

from gamut_b200.types import PixelType as PT, pixelTypeSize

pytestmark = pytest.mark.gpu
f32 = np.float32


def make_src(t, n, rng):
    """Pixels of type t with in-range values (floats in [0,1]) plus edge values."""
    comp = int(t) % 3
    ch = pixelTypeSize(t) // (1, 2, 4)[comp]
    if comp == 0:
        a = rng.integers(0, 256, n * ch, dtype=np.uint8)
        a[:ch * 4] = np.array([0, 255, 1, 254] * ch, np.uint8)[:ch * 4]
    elif comp == 1:
        a = rng.integers(0, 65536, n * ch, dtype=np.uint16)
        a[:ch * 4] = np.array([0, 65535, 1, 65534] * ch, np.uint16)[:ch * 4]
    else:
        a = rng.random(n * ch, dtype=f32)
        a[:ch * 4] = np.array([0.0, 1.0, 0.5, 1.0 / 255] * ch, f32)[:ch * 4]
    b = a.view(np.uint8).copy()
    if ch in (2, 4) and n > 16:   # a few zero alphas for the un-premultiply guard
        px = b.reshape(n, -1)
        csz = (1, 2, 4)[comp]
        px[8:12, (ch - 1) * csz:] = 0
    return b


@pytest.mark.parametrize("s", range(18))
def test_all_pairs_bit_exact(gb, oracle, s):
    rng = np.random.default_rng(100 + s)
    w, h = 1031, 3       # odd width: exercises tails and unaligned rows
    src = make_src(s, w * h, rng)
    for d in range(18):
        exp = np.zeros(w * h * pixelTypeSize(d), np.uint8)
        got = np.zeros_like(exp)
        assert oracle.scanlines_convert(s, src, w * pixelTypeSize(s), d, exp, w * pixelTypeSize(d), w, h)
        assert gb.scanlinesConvert(s, src, w * pixelTypeSize(s), d, got, w * pixelTypeSize(d), w, h), gb.last_error()
        assert np.array_equal(got, exp), f"{PT(s).name}->{PT(d).name}: {np.flatnonzero(got != exp)[:8]}"


def test_rgba8_rgbaf32_all_byte_values(gb, oracle):
    src = np.arange(256, dtype=np.uint8).repeat(4)
    exp = np.zeros(256 * 16, np.uint8); got = np.zeros_like(exp)
    oracle.scanlines_convert(PT.rgba8, src, 1024, PT.rgbaf32, exp, 4096, 256, 1)
    assert gb.scanlinesConvert(PT.rgba8, src, 1024, PT.rgbaf32, got, 4096, 256, 1)
    assert np.array_equal(got, exp)
    back = np.zeros(1024, np.uint8)
    assert gb.scanlinesConvert(PT.rgbaf32, got, 4096, PT.rgba8, back, 1024, 256, 1)
    assert np.array_equal(back, src)


def test_negative_and_padded_pitch(gb, oracle):
    w, h = 257, 9
    rng = np.random.default_rng(7)
    for s, d in [(PT.rgba8, PT.rgbaf32), (PT.rgbaf32, PT.rgba8), (PT.rgb8, PT.la16), (PT.l8, PT.rgba8), (PT.rgb16, PT.rgb16)]:
        sp = w * pixelTypeSize(s) + 8
        dp = w * pixelTypeSize(d) + 12
        src = make_src(s, (sp // pixelTypeSize(s) + 1) * h, rng)[:sp * h].copy()
        exp = np.full(dp * h, 0x5A, np.uint8); got = exp.copy()
        # read bottom-up (negative src pitch), write with a padded pitch
        assert oracle.scanlines_convert(s, src, -sp, d, exp, dp, w, h, src_off=(h - 1) * sp)
        assert gb.scanlinesConvert(s, src, -sp, d, got, dp, w, h, src_offset=(h - 1) * sp), gb.last_error()
        assert np.array_equal(got, exp)
        # negative destination pitch
        exp2 = np.full(dp * h, 0x5A, np.uint8); got2 = exp2.copy()
        assert oracle.scanlines_convert(s, src, sp, d, exp2, -dp, w, h, dst_off=(h - 1) * dp)
        assert gb.scanlinesConvert(s, src, sp, d, got2, -dp, w, h, dest_offset=(h - 1) * dp)
        assert np.array_equal(got2, exp2)


def test_empty_and_tiny(gb):
    a = np.zeros(16, np.uint8)
    assert gb.scanlinesConvert(PT.rgba8, a, 0, PT.rgbaf32, a, 0, 0, 0)
    assert gb.scanlinesConvert(PT.rgba8, a, 4, PT.rgbaf32, a, 16, 0, 1)
    assert not gb.scanlinesConvert(99, a, 4, PT.rgbaf32, a, 16, 1, 1)


def test_config2_full_size_properties(gb, oracle):
    """BASELINE config 2: 8192x8192 rgba8 <-> rgbaf32 on the device-resident path.
    Full-size check by properties: u8 -> f32 -> u8 is the identity, and the f32 image equals a
    256-entry LUT gather (the LUT itself is checked against the oracle)."""
    import torch
    from gamut_b200 import scanlinesConvertDevice
    W = H = 8192
    g = torch.Generator(device="cuda").manual_seed(1)
    src = torch.randint(0, 256, (H, W, 4), dtype=torch.uint8, device="cuda", generator=g)
    f = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    back = torch.empty_like(src)
    st = torch.cuda.current_stream().cuda_stream
    assert scanlinesConvertDevice(PT.rgba8, src.data_ptr(), W * 4, PT.rgbaf32, f.data_ptr(), W * 16, W, H, st)
    assert scanlinesConvertDevice(PT.rgbaf32, f.data_ptr(), W * 16, PT.rgba8, back.data_ptr(), W * 4, W, H, st)
    torch.cuda.synchronize()
    assert torch.equal(back, src)
    lut_src = np.arange(256, dtype=np.uint8).repeat(4)
    lut = np.zeros(256 * 16, np.uint8)
    oracle.scanlines_convert(PT.rgba8, lut_src, 1024, PT.rgbaf32, lut, 4096, 256, 1)
    lut_t = torch.from_numpy(lut.view(f32).reshape(256, 4)[:, 0].copy()).cuda()
    assert torch.equal(f, lut_t[src.long()])
    # a 64-row band against the oracle, byte for byte
    band = src[1000:1064].cpu().numpy().reshape(-1)
    exp = np.zeros(64 * W * 16, np.uint8)
    oracle.scanlines_convert(PT.rgba8, band, W * 4, PT.rgbaf32, exp, W * 16, W, 64)
    assert np.array_equal(f[1000:1064].cpu().numpy().reshape(-1).view(np.uint8), exp)


def test_reference_text_vectors(gb):
    """The CUDA converters against vectors generated from the reference's source text (tests/golden/
    gen_from_reference.py: the 46 scanline_convert_* bodies of scanline.d executed with IEEE-single semantics) --
    no oracle in between. Single-stage pairs: the source or the destination is the intermediate type."""
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_scanline.npz"))
    w = int(d["width"])
    n = 0
    for name in [str(x) for x in d["names"]]:
        src_t, dst_t = name.split("_to_")
        if src_t not in PT.__members__ or dst_t not in PT.__members__ or (src_t, dst_t) == ("l8", "rgb8"):
            continue                     # bgra8 / bgr8 / l8->rgb8 helpers are outside scanlinesConvert (scanline.d:139,812,826)
        x = np.ascontiguousarray(d["in_" + name]).view(np.uint8)
        exp = np.ascontiguousarray(d["out_" + name]).view(np.uint8)
        got = np.zeros_like(exp)
        assert gb.scanlinesConvert(PT[src_t], x, x.size, PT[dst_t], got, got.size, w, 1), name
        assert np.array_equal(got, exp), name
        n += 1
    assert n >= 40
