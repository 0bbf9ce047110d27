"""GPU parity: CUDA QOI / LZ4 / QOI-Plane10 (through the C ABI) vs the CPU oracle, bit-exact.
Reference: codecs/qoi.d, codecs/lz4.d, codecs/qoiplane10.d, plugins/qoix.d."""
import numpy as np
import pytest

from qoixutil import depth_map_la, qoi_bytes, qoi_test_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codecs(gb):
    from gamut_b200 import codecs
    return codecs


def check_qoix(codecs, oracle, data, flags=0):
    exp = oracle.qoix_decode(data, flags)
    got = codecs.qoix_decode(data, flags)
    if exp is None:
        assert got is None
        return None
    assert got is not None
    assert np.array_equal(got[0], exp[0]) and got[2] == exp[2]
    for f in ("width", "height", "pitchBytes", "channels", "bitdepth", "colorspace", "compression"):
        assert getattr(got[1], f) == getattr(exp[1], f), f
    assert got[1].pixelAspectRatio == exp[1].pixelAspectRatio and got[1].resolutionY == exp[1].resolutionY
    return got


@pytest.mark.parametrize("c", [3, 4])
def test_qoi(codecs, oracle, c):
    for (h, w) in [(64, 80), (1, 1), (7, 300)]:
        data = qoi_bytes(qoi_test_image(h, w, c, h))
        for ch in (0, 3, 4):
            exp, ed = oracle.qoi_decode(data, ch)
            got, gd = codecs.qoi_decode(data, ch)
            assert np.array_equal(got, exp) and (gd.width, gd.height, gd.channels, gd.colorspace) == (ed.width, ed.height, ed.channels, ed.colorspace)
    assert codecs.qoi_decode(data, 2) is None and codecs.qoi_decode(data[:20], 0) is None


def test_config1_qoi_512(codecs, oracle):
    """BASELINE config 1: one 512x512 RGBA8 QOI image (gradient + noise + flat rectangles + alpha ramp)."""
    img = qoi_test_image(512, 512, 4, 1234)
    data = qoi_bytes(img)
    got, _ = codecs.qoi_decode(data, 0)
    assert np.array_equal(got, img) and np.array_equal(got, oracle.qoi_decode(data, 0)[0])


@pytest.mark.parametrize("c", [1, 2])
def test_plane10(codecs, oracle, c):
    for (h, w) in [(33, 47), (1, 1), (2, 300), (64, 64), (200, 333)]:
        img = depth_map_la(h, w, 3, c)
        for force in (False, True):
            data = oracle.qoix_encode(img, 10, force_lz4=force, par=1.5, dpi=96.0)
            got = check_qoix(codecs, oracle, data)
            assert np.array_equal(got[0], img)


def test_plane10_noise_and_lz4_matches(codecs, oracle):
    rng = np.random.default_rng(1)
    v = rng.integers(0, 1024, (50, 70, 2))
    img = ((v << 6) | (v >> 4)).astype(np.uint16)                 # worst case: DIFF4 / LA everywhere
    check_qoix(codecs, oracle, oracle.qoix_encode(img, 10, force_lz4=True))
    flat = np.zeros((300, 300, 2), np.uint16) + 0x8020            # long runs; LZ4 with long overlapping matches
    check_qoix(codecs, oracle, oracle.qoix_encode(flat, 10, force_lz4=True))
    stripes = np.zeros((128, 256, 1), np.int64)
    stripes[:, ::3] = 700
    check_qoix(codecs, oracle, oracle.qoix_encode(((stripes << 6) | (stripes >> 4)).astype(np.uint16), 10, force_lz4=True))


def test_rejects_and_corrupt(codecs, oracle):
    img = depth_map_la(20, 30, 1, 2)
    enc = oracle.qoix_encode(img, 10, force_lz4=True)
    for mut in ((15, 2), (12, 3), (14, 9), (13, 7), (16, 5)):
        bad = bytearray(enc); bad[mut[0]] = mut[1]
        check_qoix(codecs, oracle, bytes(bad))
    check_qoix(codecs, oracle, enc[:20])
    check_qoix(codecs, oracle, enc[:40])                          # truncated LZ4 block
    check_qoix(codecs, oracle, enc, 0x10000 | 0x80000)
    raw = oracle.qoix_encode(img, 10)
    check_qoix(codecs, oracle, raw[:25] + b"\xff" * 8)          # early END -> zero pixels
    check_qoix(codecs, oracle, raw[:60])                          # truncated opcode stream
    assert check_qoix(codecs, oracle, enc) is not None


def test_batch(codecs, oracle):
    files = [oracle.qoix_encode(depth_map_la(40, 50, 1, 2), 10, force_lz4=True), b"junk" * 10,
             oracle.qoix_encode(depth_map_la(17, 90, 2, 1), 10), oracle.qoix_encode(depth_map_la(64, 64, 3, 2), 10, force_lz4=True)[:100]]
    b = codecs.qoix_decode_batch(files, 0)
    try:
        for i, f in enumerate(files):
            exp = oracle.qoix_decode(f, 0)
            got = b.to_host(i)
            if exp is None:
                assert got is None
            else:
                assert np.array_equal(got, exp[0]) and b.images[i].pixel_type == exp[2]
    finally:
        b.free()


def test_config5_shape_2048(codecs, oracle):
    """BASELINE config 5 shape: 2048x2048 10-bit LA + LZ4, one image exact."""
    img = depth_map_la(2048, 2048, 5, 2)
    data = oracle.qoix_encode(img, 10, force_lz4=True)
    got = check_qoix(codecs, oracle, data)
    assert np.array_equal(got[0], img)
