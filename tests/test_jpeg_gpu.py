"""GPU parity: CUDA baseline JPEG decode (through the C ABI) vs the CPU oracle. The contract allows
+-1 LSB per channel for JPEG; the kernels use the reference's integer arithmetic, so the tests demand
bit-exactness. Reference: source/gamut/codecs/jpegload.d."""
import io
import math
import os

import numpy as np
import pytest
from PIL import Image as PILImage

from jpegutil import encode, photo

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def codecs(gb):
    from gamut_b200 import codecs
    return codecs


def same(a, b):
    return a == b or (math.isnan(a) and math.isnan(b))


def check(codecs, oracle, data, req=-1):
    exp = oracle.jpeg_load(data, req)
    got = codecs.jpeg_load(data, req)
    if exp is None:
        assert got is None
        return None
    assert got is not None, "CUDA decode failed where the oracle succeeded"
    px, ac, par, dpi = exp
    assert got.pixels.shape == px.shape
    d = np.abs(got.pixels.astype(int) - px.astype(int))
    assert d.max() == 0, f"max diff {d.max()} at {np.argwhere(d > 0)[:4]}"
    assert got.actual_comps == ac and same(got.pixelAspectRatio, par) and same(got.dotsPerInchY, dpi)
    return got


def test_reference_fixtures(codecs, oracle):
    data = open(os.path.join(G, "issue35.jpg"), "rb").read()
    for req in (-1, 1, 3, 4):
        assert check(codecs, oracle, data, req) is not None
    check(codecs, oracle, b"")                                   # issue46.jpg: must fail ...
    assert check(codecs, oracle, data) is not None               # ... and not poison the next load
    check(codecs, oracle, open(os.path.join(G, "issue76.png"), "rb").read())
    check(codecs, oracle, data, 2)


@pytest.mark.parametrize("ss", [0, 1, 2])
@pytest.mark.parametrize("q", [35, 90, 100])
def test_subsampling_and_quality(codecs, oracle, ss, q):
    for (h, w) in [(150, 203), (16, 16), (1, 1), (17, 33), (240, 320)]:
        img = photo(h, w, 3, 7 * ss + q + h)
        for req in (-1, 1, 4):
            check(codecs, oracle, encode(img, q, ss), req)


def test_h1v2(codecs, oracle):
    # 4:4:0 (h1v2) is not offered by PIL's keyword; transpose trick: libjpeg writes 2x1 for subsampling=1,
    # so build an h1v2 stream by patching the SOF sampling factors of a 4:4:4 stream is not valid data;
    # use OpenCV's explicit sampling-factor flag instead when available.
    cv2 = pytest.importorskip("cv2")
    flag = getattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR", None)
    val = getattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR_440", None)
    if flag is None or val is None:
        pytest.skip("cv2 without sampling factor control")
    img = photo(90, 70, 3, 21)
    ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 90, flag, val])
    assert ok
    assert check(codecs, oracle, buf.tobytes()) is not None


def test_grey(codecs, oracle):
    for (h, w) in [(97, 131), (8, 8), (5, 3)]:
        g = photo(h, w, 1, 3)
        for req in (-1, 1, 3, 4):
            check(codecs, oracle, encode(g, 85), req)


def test_restart_intervals(codecs, oracle):
    img = photo(120, 176, 3, 9)
    for ss in (0, 1, 2):
        base = check(codecs, oracle, encode(img, 88, ss))
        for kw in (dict(restart_rows=1), dict(restart_blocks=3), dict(restart_blocks=1), dict(restart_rows=3)):
            r = check(codecs, oracle, encode(img, 88, ss, **kw))
            assert np.array_equal(r.pixels, base.pixels)


def test_optimized_huffman_and_density(codecs, oracle):
    img = photo(100, 100, 3, 12)
    check(codecs, oracle, encode(img, 75, 2, optimize=True))
    check(codecs, oracle, encode(img, 80, 0, dpi=(300, 150)))
    data = encode(img, 80, 0)
    i = data.index(b"\xff\xe0")
    n = int.from_bytes(data[i + 2:i + 4], "big")
    r = check(codecs, oracle, data[:i] + data[i + 2 + n:])
    assert math.isnan(r.dotsPerInchY)


def test_failures(codecs, oracle):
    img = photo(40, 40, 3, 2)
    data = encode(img, 90, 2)
    check(codecs, oracle, data[:200])
    check(codecs, oracle, b"\xff\xd8\xff\xd9")


def test_progressive(codecs, oracle):
    """Progressive files (SOF2; init_progressive / decode_scan / load_next_row, jpegload.d:3299-3683, :2259-2332):
    bit-exact against the oracle for every subsampling, grey, optimised tables, restart intervals inside the scans,
    every req_comps; the probe classifies them as decodable."""
    for (h, w, c, ss, q, kw) in ((48, 64, 3, 2, 90, {}), (77, 130, 3, 0, 75, {}), (150, 200, 3, 1, 85, {}), (61, 97, 1, 0, 90, {}),
                                  (250, 333, 3, 2, 95, {"optimize": True}), (256, 256, 3, 2, 50, {"restart_blocks": 7}),
                                  (173, 211, 3, 1, 80, {"restart_rows": 2}), (9, 7, 3, 2, 60, {}), (8, 8, 1, 0, 100, {})):
        data = encode(photo(h, w, c, 3 + h), q, ss, progressive=True, **kw)
        assert data.count(b"\xff\xda") > 1 and codecs.jpeg_probe(data) == 0
        check(codecs, oracle, data)
    data = encode(photo(40, 56, 3, 9), 85, 2, progressive=True)
    for rc in (1, 3, 4):
        check(codecs, oracle, data, rc)
    assert codecs.jpeg_probe(b"nope") == -1


def test_progressive_corrupt(codecs, oracle):
    """Truncated and inconsistent progressive files: same accept / reject decision and the same pixels as the oracle."""
    prog = encode(photo(40, 40, 3, 2), 90, 2, progressive=True)
    for cut in (200, len(prog) // 2, len(prog) - 3):
        check(codecs, oracle, prog[:cut])
    i = prog.rindex(b"\xff\xda")
    n = int.from_bytes(prog[i + 2:i + 4], "big")
    bad = bytearray(prog)
    bad[i + 2 + n - 1] = 0x31
    check(codecs, oracle, bytes(bad))


def test_progressive_batch_and_image(codecs, oracle, gb):
    """A batch that mixes progressive and sequential files, and Image.loadFromMemory on a progressive file."""
    from gamut_b200.image import Image
    files = [encode(photo(64, 48, 3, 1), 90, 2, progressive=True), encode(photo(64, 48, 3, 1), 90, 2),
             encode(photo(33, 65, 1, 2), 70, progressive=True), b"nope", encode(photo(120, 90, 3, 4), 80, 1, progressive=True)]
    b = codecs.jpeg_decode_batch(files, 4)
    try:
        for i, f in enumerate(files):
            exp = oracle.jpeg_load(f, 4)
            got = b.to_host(i)
            if exp is None:
                assert got is None and b.images[i].status == 0
            else:
                assert np.array_equal(got, exp[0])
    finally:
        b.free()
    im = Image()
    im.loadFromMemory(files[0])
    assert not im.isError() and im.width() == 48 and im.height() == 64


def test_batch_mixed(codecs, oracle):
    files = [encode(photo(64, 48, 3, 1), 90, 2), b"nope", encode(photo(33, 65, 1, 2), 70),
             encode(photo(80, 80, 3, 3), 95, 0, restart_rows=1), encode(photo(50, 70, 3, 4), 60, 1)]
    b = codecs.jpeg_decode_batch(files, -1)
    try:
        for i, f in enumerate(files):
            exp = oracle.jpeg_load(f, -1)
            got = b.to_host(i)
            if exp is None:
                assert got is None and b.images[i].status == 0
            else:
                assert np.array_equal(got, exp[0])
    finally:
        b.free()


def test_config4_shape_4k(codecs, oracle):
    """BASELINE config 4 shape: 3840x2160 4:2:0 quality 90, one image exact against the oracle."""
    img = photo(2160, 3840, 3, 44)
    check(codecs, oracle, encode(img, 90, 2))
