"""BMP decode and format detection on the GPU path against the oracle, through the C ABI."""
import os

import numpy as np
import pytest

from bmputil import variants, broken

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def codecs(gb):
    from gamut_b200 import codecs
    return codecs


def check(codecs, oracle, data, req=0):
    exp = oracle.bmp_load(data, req)
    got = codecs.bmp_load(data, req)
    if exp is None or exp[0].size == 0:
        assert got is None
        return None
    assert got is not None, "CUDA decode failed where the oracle succeeded"
    px, comp, ppmX, ppmY, ratio = exp
    assert got.pixels.shape == px.shape and np.array_equal(got.pixels, px)
    assert got.file_channels == comp and got.ppmX == ppmX and got.ppmY == ppmY
    assert got.pixelRatio == ratio or (np.isnan(got.pixelRatio) and np.isnan(ratio))
    return got


def test_variants_all_req(codecs, oracle):
    """Every branch of stbi__bmp_load (1/4/8-bit palettes, 16/32-bit fields, 24/32-bit easy paths, OS/2 / V3 / V4 / V5
    headers, top-down, gaps before the pixel data), every req_comp, bit-exact."""
    rng = np.random.default_rng(5)
    for name, f in variants(rng):
        for req in (0, 1, 2, 3, 4):
            assert check(codecs, oracle, f, req) is not None, (name, req)


def test_broken(codecs, oracle):
    rng = np.random.default_rng(7)
    for name, f in broken(rng):
        check(codecs, oracle, f, 0)
        check(codecs, oracle, f, 4)


def test_issue67_and_image(codecs, oracle, gb):
    """The reference's BMP KAT (examples/test-suite/source/main.d:161-170) through Image.loadFromMemory."""
    from gamut_b200.image import Image, kStrImageFormatNoLoadSupport, kStrImageFormatUnidentified
    from gamut_b200.types import ImageFormat, PixelType as PT, LOAD_RGB, LOAD_ALPHA, LOAD_GREYSCALE, LOAD_FP32, LAYOUT_VERT_FLIPPED
    data = open(os.path.join(G, "issue67.bmp"), "rb").read()
    check(codecs, oracle, data)
    im = Image()
    assert im.loadFromMemory(data)
    assert im.width() == 32 and im.height() == 32 and im.type() == PT.rgb8
    assert abs(im.dotsPerInchX() - 200) < 0.1 and abs(im.dotsPerInchY() - 100) < 0.1 and abs(im.pixelAspectRatio() - 2) < 0.01
    from oracle import pyimage
    for flags in (0, LOAD_RGB | LOAD_ALPHA, LOAD_GREYSCALE | LOAD_FP32, LOAD_RGB | LAYOUT_VERT_FLIPPED):
        exp = pyimage.load_from_memory(data, flags)
        im.loadFromMemory(data, flags)
        assert (exp.error is None) == im.isValid()
        if exp.error is None:
            assert int(im.type()) == exp.type and im.pitchInBytes() == exp.pitch
            for y in range(im.height()):
                assert np.array_equal(im.scanline(y), exp.scanline(y))
        im2 = Image()
        im2.loadFromMemoryStaged(data, flags)
        assert im2.isValid() == im.isValid()
    # detected formats without a loader, and unknown data (image.d:1756-1770)
    assert Image.identifyFormatFromMemory(data) == ImageFormat.BMP
    im.loadFromMemory(b"GIF89a" + b"\0" * 64)
    assert im.isError() and im.errorMessage() == kStrImageFormatNoLoadSupport
    im.loadFromMemory(b"nope, not an image")
    assert im.isError() and im.errorMessage() == kStrImageFormatUnidentified


def test_identify_format_matches_oracle(codecs, oracle):
    rng = np.random.default_rng(11)
    samples = [b"", b"B", b"BM", b"qoif", b"qoix", b"DDS ", b"GIF87a", b"GIF89a", b"\xff\xd8", b"\xff\x0a", b"\xa5", b"\x89PNG\r\n\x1a\n"]
    samples += [f for _, f in variants(rng)] + [f for _, f in broken(rng)]
    for _ in range(3000):                 # fuzz the TGA header test (codecs/tga.d:313-382) and the magics
        n = int(rng.integers(0, 24))
        b = bytearray(rng.integers(0, 256, n, dtype=np.uint8).tobytes())
        if n >= 3 and rng.random() < 0.7:
            b[1] = int(rng.integers(0, 3)); b[2] = int(rng.choice([1, 2, 3, 9, 10, 11, 4]))
        if n >= 17 and rng.random() < 0.7:
            b[16] = int(rng.choice([8, 15, 16, 24, 32, 7]))
        samples.append(bytes(b))
    for s in samples:
        assert codecs.identify_format(s) == oracle.identify_format(s), s[:20]


def test_batch(codecs, oracle):
    rng = np.random.default_rng(5)
    files = [f for _, f in variants(rng)][:10] + [b"nope"]
    b = codecs.bmp_decode_batch(files, 4)
    try:
        for i, f in enumerate(files):
            exp = oracle.bmp_load(f, 4)
            got = b.to_host(i)
            if exp is None:
                assert got is None and b.images[i].status == 0
            else:
                assert np.array_equal(got, exp[0])
    finally:
        b.free()
