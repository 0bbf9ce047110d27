"""GPU parity of the TGA decoder: gb200_tga_load / gb200_tga_decode_batch / Image.loadFromMemory against
the oracle's restatement of TGADecoder (codecs/tga.d:313-588, pinned to PIL and to the format's formulas in
tests/test_oracle_tga.py) on every variant of tests/tgautil.py: grey, grey + alpha, 15/16/24/32-bit colour, palettes
with 8- and 16-bit indices, raw and run-length coded, both orientations, ID field, truncated files."""
import numpy as np
import pytest

from tgautil import make_tga, pil_tga
from test_oracle_tga import CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codecs(gb):
    from gamut_b200 import codecs
    return codecs


def same(a, b):
    return (a is None and b is None) or (a is not None and b is not None and a.shape == b.shape and np.array_equal(a, b))


@pytest.mark.parametrize("kind,kw", CASES)
def test_variants(codecs, oracle, kind, kw):
    rng = np.random.default_rng(7)
    for rle in (False, True):
        for top_down in (False, True):
            for (w, h) in [(1, 1), (13, 9), (131, 40), (640, 33)]:
                data, _ = make_tga(w, h, kind, rng, rle=rle, top_down=top_down, **kw)
                exp = oracle.tga_load(data)
                assert exp is not None and same(codecs.tga_load(data), exp)
            for cut in (len(data) - 1, len(data) // 2, 19, 18, 5, 0):
                assert same(codecs.tga_load(data[:cut]), oracle.tga_load(data[:cut]))


def test_pil_files_large_and_rejects(codecs, oracle):
    rng = np.random.default_rng(1)
    for c in (1, 3, 4):
        img = rng.integers(0, 4, (23, 37, c)).astype(np.uint8) * 80
        for rle in (False, True):
            assert same(codecs.tga_load(pil_tga(img, rle, bool(c & 1))), img)
    big = np.zeros((1080, 1920, 4), np.uint8)                    # runs and noise at a real size, both codings
    big[..., :3] = np.linspace(0, 255, 1920)[None, :, None].astype(np.uint8)
    big[200:500, 300:900] = rng.integers(0, 256, (300, 600, 4))
    big[..., 3] = 255
    for rle in (False, True):
        assert same(codecs.tga_load(pil_tga(big, rle)), big)
    data, _ = make_tga(17, 11, "bgr24", rng, rle=True, overrun=True)
    assert same(codecs.tga_load(data), oracle.tga_load(data))
    ok, _ = make_tga(5, 4, "pal24", rng, pal_len=9)
    for _ in range(100):                                         # header fuzz: both sides accept / reject the same files
        bad = bytearray(ok)
        for _ in range(int(rng.integers(1, 4))):
            bad[int(rng.integers(0, 18))] = int(rng.integers(0, 256))
        bad[12:16] = (5).to_bytes(2, "little") + (4).to_bytes(2, "little")
        assert same(codecs.tga_load(bytes(bad)), oracle.tga_load(bytes(bad)))


def test_batch(codecs, oracle):
    rng = np.random.default_rng(5)
    files = []
    for kind, kw in CASES:
        for rle in (False, True):
            files.append(make_tga(57, 23, kind, rng, rle=rle, top_down=bool(len(files) & 1), **kw)[0])
    files += [b"nope", files[3][:-1], files[0][:40]]
    b = codecs.tga_decode_batch(files)
    try:
        for i, f in enumerate(files):
            exp = oracle.tga_load(f)
            got = b.to_host(i)
            if exp is None:
                assert got is None and b.images[i].status == 0
            else:
                assert np.array_equal(got.reshape(exp.shape), exp)
    finally:
        b.free()


def test_image_load(codecs, oracle):
    """Image.loadFromMemory on TGA files (plugins/tga.d:45-105 -> convertTo): type, pitch and every scanline equal to the
    oracle-side restatement of the plugin for a few flag / layout combinations; the staged path agrees."""
    from gamut_b200.image import Image, kStrImageDecodingFailed
    from gamut_b200.types import (ImageFormat, PixelType as PT, LOAD_RGB, LOAD_ALPHA, LOAD_GREYSCALE, LOAD_FP32, LOAD_16BIT,
                                  LAYOUT_VERT_FLIPPED, LAYOUT_SCANLINE_ALIGNED_16, LAYOUT_BORDER_1)
    from oracle import pyimage
    rng = np.random.default_rng(2)
    for kind, kw, rle in [("bgr24", {}, True), ("bgra32", {}, False), ("l8", {}, True), ("la16", {}, False), ("pal24", {"pal_len": 99}, True), ("rgb16", {}, False)]:
        data, _ = make_tga(45, 31, kind, rng, rle=rle, **kw)
        assert Image.identifyFormatFromMemory(data) == ImageFormat.TGA
        for flags in (0, LOAD_RGB | LOAD_ALPHA, LOAD_GREYSCALE | LOAD_FP32, LOAD_16BIT | LAYOUT_VERT_FLIPPED, LAYOUT_SCANLINE_ALIGNED_16 | LAYOUT_BORDER_1):
            exp = pyimage.load_from_memory(data, flags)
            im = Image()
            im.loadFromMemory(data, flags)
            assert (exp.error is None) == im.isValid()
            if exp.error is None:
                assert int(im.type()) == exp.type and im.pitchInBytes() == exp.pitch and im.width() == 45 and im.height() == 31
                for y in range(im.height()):
                    assert np.array_equal(im.scanline(y), exp.scanline(y))
            im2 = Image()
            im2.loadFromMemoryStaged(data, flags)
            assert im2.isValid() == im.isValid()
            if im.isValid():
                assert int(im2.type()) == int(im.type())
                for y in range(im.height()):
                    assert np.array_equal(im2.scanline(y), im.scanline(y))
    im = Image()
    data, _ = make_tga(45, 31, "bgr24", rng)
    im.loadFromMemory(data[:-5])
    assert im.isError() and im.errorMessage() == kStrImageDecodingFailed
