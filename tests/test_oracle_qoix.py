"""Pin the QOI / LZ4 / QOIX oracle (oracle/qoix_oracle.c): QOI against PIL's independent codec, LZ4 against the
system liblz4 (both directions), QOIX through the reference's own round-trip property
(image.d:2112-2183; examples/qoix/source/main.d:113-121)."""
import ctypes as C

import numpy as np
import pytest

from qoixutil import depth_map_la, liblz4, qoi_bytes, qoi_test_image


@pytest.mark.parametrize("c", [3, 4])
def test_qoi_matches_pil(oracle, c):
    img = qoi_test_image(64, 80, c, 1)
    data = qoi_bytes(img)
    for ch in (0, 3, 4):
        px, desc = oracle.qoi_decode(data, ch)
        assert desc.channels == c and (desc.width, desc.height) == (80, 64)
        n = ch or c
        exp = img if n == c else (img[:, :, :3] if n == 3 else np.dstack([img, np.full(img.shape[:2], 255, np.uint8)]))
        assert np.array_equal(px, exp)
    assert oracle.qoi_decode(data, 2) is None and oracle.qoi_decode(data[:20], 0) is None
    assert oracle.qoi_decode(b"qoix" + data[4:], 0) is None


def test_qoi_3x1_kat(oracle):
    # image.d:2112-2183: 3x1 rgb8 [255,0,0, 15,64,255, 0,255,255] must survive encode -> decode
    img = np.array([[[255, 0, 0], [15, 64, 255], [0, 255, 255]]], np.uint8)
    px, _ = oracle.qoi_decode(qoi_bytes(img), 0)
    assert np.array_equal(px, img)


def test_lz4_against_system_liblz4(oracle):
    L = liblz4()
    if L is None:
        pytest.skip("no liblz4 runtime")
    rng = np.random.default_rng(0)
    for src in (bytes(rng.integers(0, 4, 70000, dtype=np.uint8)), b"abc" * 5000, bytes(rng.integers(0, 256, 3000, dtype=np.uint8)),
                b"x" * 100000, b"short", bytes(13)):
        # (a) a stream from the real encoder decodes identically in the oracle
        cap = L.LZ4_compressBound(len(src))
        buf = C.create_string_buffer(cap)
        n = L.LZ4_compress_default(src, buf, len(src), cap)
        assert n > 0
        out = oracle.lz4_decompress(buf.raw[:n], len(src))
        assert out is not None and out.tobytes() == src
        # (b) the oracle's generator emits valid LZ4 that the real decoder accepts
        comp = oracle.lz4_compress(src)
        dst = C.create_string_buffer(len(src) + 1)
        assert L.LZ4_decompress_safe(comp, dst, len(comp), len(src)) == len(src) and dst.raw[:len(src)] == src
    # corrupt / truncated input is rejected, not read out of bounds
    comp = oracle.lz4_compress(b"abc" * 5000)
    assert oracle.lz4_decompress(comp[:10], 15000) is None
    assert oracle.lz4_decompress(comp, 14999) is None


@pytest.mark.parametrize("c", [1, 2])
def test_plane10_roundtrip(oracle, c):
    for (h, w) in [(33, 47), (1, 1), (2, 300), (64, 64)]:
        img = depth_map_la(h, w, 3, c)
        for force in (False, True):
            enc = oracle.qoix_encode(img, 10, force_lz4=force, par=1.5, dpi=96.0)
            px, desc, t = oracle.qoix_decode(enc, 0)
            assert np.array_equal(px, img) and (desc.width, desc.height, desc.channels, desc.bitdepth) == (w, h, c, 10)
            assert t == (1 if c == 1 else 4)                     # l16 / la16 (plugins/qoix.d:476-507)
            assert desc.pixelAspectRatio == 1.5 and desc.resolutionY == 96.0
            if force:
                assert enc[16] == 1


def test_plane10_3x1_kat_and_rejects(oracle):
    img = np.array([[[1023, 0], [15, 64], [0, 1023]]], np.int64)
    img16 = ((img << 6) | (img >> 4)).astype(np.uint16)
    enc = oracle.qoix_encode(img16, 10)
    assert np.array_equal(oracle.qoix_decode(enc, 0)[0], img16)
    bad = bytearray(enc); bad[15] = 2                            # premultiplied streams are rejected (qoiplane10.d:341)
    assert oracle.qoix_decode(bytes(bad), 0) is None
    bad = bytearray(enc); bad[12] = 3
    assert oracle.qoix_decode(bytes(bad), 0) is None
    assert oracle.qoix_decode(enc[:20], 0) is None
    assert oracle.qoix_decode(enc, 0x10000 | 0x80000) is None    # invalid LoadFlags
    # early END: remaining pixels are zero in the restatement
    cut = enc[:25] + b"\xff" * 8
    px = oracle.qoix_decode(cut, 0)[0]
    assert (px == 0).all()
