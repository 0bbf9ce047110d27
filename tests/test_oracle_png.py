"""Pin the PNG oracle (oracle/png_oracle.c) against the reference's own fixtures / KATs
(examples/test-suite/source/main.d) and cross-check it against PIL on generated files."""
import io
import os
import zlib

import numpy as np
import pytest
from PIL import Image as PILImage

from pngwriter import write_png

G = os.path.join(os.path.dirname(__file__), "golden")


def rd(name):
    return open(os.path.join(G, name), "rb").read()


def test_issue76_kat(oracle):
    # examples/test-suite/source/main.d:172-190: 2x2 16-bit grey, l16 [[1875,65535],[0,2807]]
    data = rd("issue76.png")
    assert oracle.png_is16(data)
    px, info = oracle.png_load(data, 0, 1)
    assert (info.width, info.height, info.channels) == (2, 2, 1)
    assert px[:, :, 0].tolist() == [[1875, 65535], [0, 2807]]


def test_buggy_miniz_chunk_kat(oracle):
    # main.d:57-69: inflates to 594825+272 bytes with a 594825 size hint (buffer grows once)
    data = rd("buggy-miniz-chunk.bin")
    out = oracle.zlib_decode(data, 594825, 1)
    assert out is not None and out.size == 594825 + 272
    assert out.tobytes() == zlib.decompress(data)


@pytest.mark.parametrize("name", ["vst3-compatible.png", "issue65.png", "issue92-truncated-in-CRC.png",
                                  "issue92-no-IEND.png"])
def test_must_load_matches_pil(oracle, name):
    data = rd(name)
    px, info = oracle.png_load(data, 0, 0)
    assert px is not None
    ref = np.asarray(PILImage.open(io.BytesIO(data)).convert("RGBA"))
    assert info.channels == 4 and np.array_equal(px, ref)


@pytest.mark.parametrize("name", ["issue51cgbi.png", "issue51cgbi2.png"])
def test_cgbi_loads_as_stored(oracle, name):
    # main.d:72-83: loads; this port has no BGR swap / un-premultiply: pixels stay as stored
    px, info = oracle.png_load(rd(name), 0, 0)
    assert px is not None and info.channels == 4 and px.shape[0] == px.shape[1]


def test_empty_and_garbage_fail(oracle):
    assert oracle.png_load(b"", 0, 0)[0] is None
    assert oracle.png_load(rd("issue35.jpg"), 0, 0)[0] is None
    good = rd("issue76.png")
    assert oracle.png_load(good[:60], 0, 0)[0] is None


def synth(h, w, c, depth, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = (np.sin(xx / 7.0)[..., None] + np.cos(yy / 5.0)[..., None] + np.arange(c)[None, None, :]) * 0.2 + 0.5
    v = np.clip(base + rng.normal(0, 0.02, (h, w, c)), 0, 1)
    return (v * ((1 << depth) - 1) + 0.5).astype(np.int64)


@pytest.mark.parametrize("color,c", [(0, 1), (2, 3), (4, 2), (6, 4)])
@pytest.mark.parametrize("depth", [8, 16])
@pytest.mark.parametrize("filt", [0, 1, 2, 3, 4, (0, 1, 2, 3, 4, 4, 3, 1)])
def test_filters_vs_pil(oracle, color, c, depth, filt):
    img = synth(37, 53, c, depth, 11 * color + depth)
    data = write_png(img, color, depth, filters=filt)
    px, info = oracle.png_load(data, 0, 1 if depth == 16 else 0)
    assert px is not None and np.array_equal(px.astype(np.int64), img)
    if depth == 8:
        ref = np.asarray(PILImage.open(io.BytesIO(data)))
        assert np.array_equal(px.reshape(ref.shape), ref)


@pytest.mark.parametrize("depth", [1, 2, 4, 8])
def test_low_depth_grey_and_palette(oracle, depth):
    h, w = 19, 29
    img = synth(h, w, 1, depth, depth)
    scale = {1: 255, 2: 85, 4: 17, 8: 1}[depth]
    px, _ = oracle.png_load(write_png(img, 0, depth, filters=(0, 1, 2, 3, 4)), 0, 0)
    assert np.array_equal(px[:, :, 0], (img[:, :, 0] * scale).astype(np.uint8))
    pal = np.random.default_rng(depth).integers(0, 256, (1 << depth, 3))
    data = write_png(img, 3, depth, filters=4, palette=pal)
    px, info = oracle.png_load(data, 0, 0)
    assert info.channels == 3 and np.array_equal(px, pal[img[:, :, 0]].astype(np.uint8))
    ref = np.asarray(PILImage.open(io.BytesIO(data)).convert("RGB"))
    assert np.array_equal(px, ref)
    # palette + tRNS -> 4 channels
    tr = bytes(range(0, 1 << depth))[: 1 << depth]
    px, info = oracle.png_load(write_png(img, 3, depth, filters=2, palette=pal, trns=tr), 0, 0)
    assert info.channels == 4 and np.array_equal(px[:, :, 3], np.array(list(tr), np.uint8)[img[:, :, 0]])


def test_trns_colour_key(oracle):
    img = synth(16, 16, 3, 8, 5)
    key = img[3, 4]
    tr = b"".join(int(v).to_bytes(2, "big") for v in key)
    px, info = oracle.png_load(write_png(img, 2, 8, filters=1, trns=tr), 0, 0)
    assert info.channels == 4
    m = (img == key).all(axis=2)
    assert np.array_equal(px[:, :, 3], np.where(m, 0, 255).astype(np.uint8)) and np.array_equal(px[:, :, :3], img.astype(np.uint8))
    g = synth(9, 9, 1, 16, 6)
    tr = int(g[2, 2, 0]).to_bytes(2, "big")
    px, info = oracle.png_load(write_png(g, 0, 16, filters=3, trns=tr), 0, 1)
    assert info.channels == 2 and np.array_equal(px[:, :, 1], np.where(g[:, :, 0] == g[2, 2, 0], 0, 65535))


@pytest.mark.parametrize("color,c,depth", [(6, 4, 8), (2, 3, 16), (0, 1, 2), (3, 1, 4), (4, 2, 8)])
def test_adam7(oracle, color, c, depth):
    h, w = 21, 13
    img = synth(h, w, c, depth, 77)
    pal = np.random.default_rng(1).integers(0, 256, (1 << depth, 3)) if color == 3 else None
    data = write_png(img, color, depth, filters=(4, 3, 2, 1, 0), interlace=True, palette=pal)
    px, info = oracle.png_load(data, 0, 1 if depth == 16 else 0)
    assert px is not None
    if color == 3:
        exp = pal[img[:, :, 0]]
    elif depth < 8:
        exp = img * {1: 255, 2: 85, 4: 17}[depth]
    else:
        exp = img
    assert np.array_equal(px.astype(np.int64), exp)


def test_req_comp_and_depth_conversion(oracle):
    img = synth(11, 17, 4, 8, 9)
    data = write_png(img, 6, 8, filters=4)
    r, g, b, a = [img[:, :, k] for k in range(4)]
    y = ((r * 77 + g * 150 + b * 29) >> 8)
    px, info = oracle.png_load(data, 1, 0)
    assert info.file_channels == 4 and np.array_equal(px[:, :, 0], y)
    px, _ = oracle.png_load(data, 2, 0)
    assert np.array_equal(px[:, :, 0], y) and np.array_equal(px[:, :, 1], a)
    px, _ = oracle.png_load(data, 3, 1)                       # 8 -> 16: v*257 (stbdec.d:662)
    assert np.array_equal(px.astype(np.int64), img[:, :, :3] * 257)
    g16 = synth(7, 5, 1, 16, 3)
    px, _ = oracle.png_load(write_png(g16, 0, 16, filters=2), 4, 0)   # 16 -> 8: >>8 (stbdec.d:645)
    assert np.array_equal(px[:, :, 0], g16[:, :, 0] >> 8) and (px[:, :, 3] == 255).all()
    rgb = synth(6, 6, 3, 8, 4)
    px, _ = oracle.png_load(write_png(rgb, 2, 8, filters=3), 4, 0)    # alpha inserted during unfilter
    assert np.array_equal(px[:, :, :3], rgb) and (px[:, :, 3] == 255).all()


def test_phys_and_multi_idat_and_trailing(oracle):
    img = synth(40, 40, 3, 8, 2)
    data = write_png(img, 2, 8, filters=4, idat_split=97, phys=(3780, 3780, 1), extra_raw=b"\0" * 300)
    px, info = oracle.png_load(data, 0, 0)
    assert np.array_equal(px, img) and info.ppmX == 3780 and info.pixelRatio == 1.0
    px, info = oracle.png_load(write_png(img, 2, 8, phys=(2, 1, 0)), 0, 0)
    assert info.ppmX == -1 and info.pixelRatio == 2.0
    assert oracle.png_load(write_png(img, 2, 8, iend=False), 0, 0)[0] is not None     # issue #92
