"""GPU parity: CUDA PNG decode (inflate + unfilter + finish kernels, through the C ABI) vs the CPU
oracle and the reference's fixtures, bit-exact. Reference: source/gamut/codecs/stbdec.d:1260-2110."""
import os
import zlib

import numpy as np
import pytest

from pngwriter import write_png
from test_oracle_png import synth, rd

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codecs(gb):
    from gamut_b200 import codecs
    return codecs


def check(codecs, oracle, data, req_comp=0, want16=0):
    exp, info = oracle.png_load(data, req_comp, want16)
    got = codecs.png_load(data, req_comp, bool(want16))
    if exp is None:
        assert got is None
        return None
    assert got is not None, "CUDA decode failed where the oracle succeeded"
    assert got.pixels.shape == exp.shape and got.pixels.dtype == exp.dtype
    assert np.array_equal(got.pixels, exp), np.argwhere(got.pixels != exp)[:5]
    assert got.file_channels == info.file_channels
    assert (got.ppmX, got.ppmY) == (info.ppmX, info.ppmY)
    assert got.pixelRatio == info.pixelRatio or (np.isnan(got.pixelRatio) and np.isnan(info.pixelRatio))
    return got


def test_issue76_kat(codecs):
    # examples/test-suite/source/main.d:172-190
    data = rd("issue76.png")
    assert codecs.png_is16(data)
    r = codecs.png_load(data, 0, True)
    assert r.pixels[:, :, 0].tolist() == [[1875, 65535], [0, 2807]]


@pytest.mark.parametrize("name", ["vst3-compatible.png", "issue65.png", "issue92-truncated-in-CRC.png",
                                  "issue92-no-IEND.png", "issue51cgbi.png", "issue51cgbi2.png", "issue76.png"])
@pytest.mark.parametrize("req,w16", [(0, 0), (3, 0), (2, 1)])
def test_reference_fixtures(codecs, oracle, name, req, w16):
    assert check(codecs, oracle, rd(name), req, w16) is not None


def test_must_fail(codecs, oracle):
    for data in (b"", rd("issue35.jpg"), rd("issue76.png")[:60], rd("issue65.png")[:200000]):
        check(codecs, oracle, data)
    # corrupt filter byte and corrupt deflate stream
    img = synth(8, 8, 3, 8, 1)
    good = write_png(img, 2, 8, filters=1)
    assert check(codecs, oracle, good) is not None
    bad = bytearray(write_png(img, 2, 8, filters=0, level=0))
    i = bad.index(b"IDAT") + 4 + 2 + 5       # first filter byte of a stored block
    bad[i] = 9
    check(codecs, oracle, bytes(bad))
    bad2 = bytearray(write_png(synth(64, 64, 3, 8, 2), 2, 8, filters=4))
    j = bad2.index(b"IDAT") + 40
    bad2[j] ^= 0xFF
    check(codecs, oracle, bytes(bad2))
    # a failure must not poison the next load (main.d:38-49)
    assert check(codecs, oracle, good) is not None


@pytest.mark.parametrize("color,c", [(0, 1), (2, 3), (4, 2), (6, 4)])
@pytest.mark.parametrize("depth", [8, 16])
def test_filters(codecs, oracle, color, c, depth):
    for filt in [0, 1, 2, 3, 4, (0, 1, 2, 3, 4, 4, 3, 1)]:
        img = synth(67, 131, c, depth, 11 * color + depth)
        data = write_png(img, color, depth, filters=filt)
        check(codecs, oracle, data, 0, 1 if depth == 16 else 0)
    check(codecs, oracle, data, (c % 4) + 1, 0)


@pytest.mark.parametrize("depth", [1, 2, 4, 8])
def test_low_depth_palette_trns(codecs, oracle, depth):
    img = synth(19, 29, 1, depth, depth)
    check(codecs, oracle, write_png(img, 0, depth, filters=(0, 1, 2, 3, 4)))
    pal = np.random.default_rng(depth).integers(0, 256, (1 << depth, 3))
    for req in (0, 1, 2, 3, 4):
        check(codecs, oracle, write_png(img, 3, depth, filters=4, palette=pal), req)
        check(codecs, oracle, write_png(img, 3, depth, filters=2, palette=pal, trns=bytes(range(1 << depth))[: 1 << depth]), req)
    check(codecs, oracle, write_png(img, 0, depth, filters=3, trns=int(img[1, 1, 0]).to_bytes(2, "big")), 0)
    check(codecs, oracle, write_png(img, 0, depth, filters=3, trns=int(img[1, 1, 0]).to_bytes(2, "big")), 4, 1)


def test_trns_rgb(codecs, oracle):
    img = synth(16, 16, 3, 8, 5)
    tr = b"".join(int(v).to_bytes(2, "big") for v in img[3, 4])
    for req in (0, 1, 2, 3, 4):
        check(codecs, oracle, write_png(img, 2, 8, filters=1, trns=tr), req)
    g = synth(9, 9, 3, 16, 6)
    tr = b"".join(int(v).to_bytes(2, "big") for v in g[2, 2])
    check(codecs, oracle, write_png(g, 2, 16, filters=3, trns=tr), 0, 1)
    check(codecs, oracle, write_png(g, 2, 16, filters=3, trns=tr), 0, 0)


@pytest.mark.parametrize("color,c,depth", [(6, 4, 8), (2, 3, 16), (0, 1, 2), (3, 1, 4), (4, 2, 8), (0, 1, 1)])
def test_adam7(codecs, oracle, color, c, depth):
    for (h, w) in [(21, 13), (1, 1), (3, 9), (64, 64)]:
        img = synth(h, w, c, depth, 77)
        pal = np.random.default_rng(1).integers(0, 256, (1 << depth, 3)) if color == 3 else None
        data = write_png(img, color, depth, filters=(4, 3, 2, 1, 0), interlace=True, palette=pal)
        check(codecs, oracle, data, 0, 1 if depth == 16 else 0)
        check(codecs, oracle, data, 4, 0)


def test_deflate_block_types_and_sizes(codecs, oracle):
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, (96, 160, 3))
    flat = np.zeros((300, 500, 4), np.int64) + 7
    for img, color in ((noise, 2), (flat, 6), (synth(200, 333, 4, 8, 8), 6)):
        for level in (0, 1, 6, 9):
            check(codecs, oracle, write_png(img, color, 8, filters=(1, 4, 2), level=level, idat_split=4093))
    # fixed-Huffman blocks (Z_FIXED) and trailing bytes after the image rows
    img = synth(50, 50, 3, 8, 1)
    from pngwriter import filter_rows, pack_samples, _chunk
    import struct
    raw = filter_rows(pack_samples(img, 8), 3, 4) + b"\0" * 100
    co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)
    comp = co.compress(raw) + co.flush()
    data = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", 50, 50, 8, 2, 0, 0, 0)) + _chunk(b"IDAT", comp) + _chunk(b"IEND", b"")
    check(codecs, oracle, data)


def test_batch_mixed_with_failures(codecs, oracle):
    files = [rd("issue65.png"), b"junk", write_png(synth(33, 17, 2, 16, 1), 4, 16, filters=4), rd("vst3-compatible.png"),
             rd("issue76.png")[:80], write_png(synth(40, 40, 1, 4, 2), 0, 4, filters=2, interlace=True)]
    b = codecs.png_decode_batch(files, 0, -1)
    try:
        for i, f in enumerate(files):
            w16 = 1 if oracle.png_is16(f) else 0
            exp, info = oracle.png_load(f, 0, w16)
            got = b.to_host(i)
            if exp is None:
                assert got is None and b.images[i].status == 0
            else:
                assert b.images[i].status == 1 and np.array_equal(got, exp)
    finally:
        b.free()


def test_config3_shape_1080p_rgba(codecs, oracle):
    """BASELINE config 3 shape: 1920x1080 RGBA8, adaptive filters (PIL-like mix), one image exact."""
    img = synth(1080, 1920, 4, 8, 42)
    data = write_png(img, 6, 8, filters=(4, 4, 3, 1, 2, 4, 0, 3), level=6, idat_split=65536)
    check(codecs, oracle, data)


def test_rgba8_wavefront_geometry(codecs, oracle):
    """Widths/heights around every boundary of the 4-byte-pixel wavefront kernel (16-column chunks, 32-row bands,
    31 steps of skew, misaligned row starts): RGBA8 and LA16 with every filter mix, bit-exact against the oracle."""
    rng = np.random.default_rng(77)
    shapes = [(1, 1), (1, 40), (2, 15), (3, 16), (5, 17), (31, 31), (32, 32), (33, 33), (34, 47), (40, 48), (63, 49),
              (64, 64), (65, 65), (97, 95), (100, 129), (130, 481), (37, 1000)]
    mixes = [4, 3, (4, 3), (0, 1, 2, 3, 4), (2, 4, 4, 1, 3, 0, 4), (4, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
             0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3)]
    for (h, w) in shapes:
        img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        img[:, :, 3] = (img[:, :, 3] // 64) * 64 + 63            # long flat runs in one channel
        filt = mixes[(h * 7 + w) % len(mixes)]
        check(codecs, oracle, write_png(img, 6, 8, filters=filt))
        la = rng.integers(0, 65536, (h, w, 2)).astype(np.uint16)
        check(codecs, oracle, write_png(la, 4, 16, filters=mixes[(h + w) % len(mixes)]), 0, 1)
