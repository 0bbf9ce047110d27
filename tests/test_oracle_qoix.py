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


@pytest.mark.parametrize("c", [3, 4])
def test_qoi_encoder_matches_pil_writer(oracle, c):
    """or_qoi_encode (qoi.d:295-426) against PIL's independent QOI writer, byte for byte: the opcode choice of the
    format's reference encoder is deterministic (run, index, diff, luma, rgb / rgba in that order), so two faithful
    encoders agree on the whole stream. PIL writes colorspace from its own default: read it from its header."""
    rng = np.random.default_rng(c)
    imgs = [qoi_test_image(64, 80, c, 1), qoi_test_image(3, 200, c, 2), np.zeros((5, 200, c), np.uint8),
            rng.integers(0, 4, (40, 40, c)).astype(np.uint8) * 70, rng.integers(0, 256, (30, 50, c)).astype(np.uint8),
            (np.cumsum(rng.integers(-9, 10, (40, 60, c)), axis=1) % 256).astype(np.uint8)]
    for img in imgs:
        pil = qoi_bytes(img)
        assert oracle.qoi_encode(img, colorspace=pil[13]) == pil
    # pitch: padded rows and a negative pitch address the same pixels
    img = imgs[0]
    wide = np.zeros((64, 100, c), np.uint8) + 77
    wide[:, :80] = img
    assert oracle.qoi_encode(wide, pitch=100 * c, shape=(64, 80, c)) == oracle.qoi_encode(img)
    # qoi_encode's refusals (qoi.d:303-311)
    assert oracle.qoi_encode(img, colorspace=2) is None
    assert oracle.qoi_encode(img, shape=(64, 80, 2)) is None and oracle.qoi_encode(img, shape=(0, 80, c)) is None


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


# ---- QOIX sub-codecs without a restated reference encoder (qoi2avg.d, qoiplane.d, qoi10b.d) ----
def _rand_img(h, w, c, seed, hi=256):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, hi, (h, w, c))
    img[h // 3: h // 2, :, :] = img[h // 3, 0, :]          # flat band
    if c >= 3:
        img[:, w // 2:, 1] = img[:, w // 2:, 0]
        img[:, w // 2:, 2] = img[:, w // 2:, 0]            # grey half
    if c in (2, 4):
        img[: h // 2, :, c - 1] = hi - 1                     # constant alpha on top
    return img


@pytest.mark.parametrize("c", [3, 4])
def test_qoi2avg_literal_roundtrip_and_fuzz(oracle, c):
    import qoixsynth as qs
    for (h, w) in [(1, 1), (9, 13), (40, 31)]:
        img = _rand_img(h, w, c, h * w + c).astype(np.uint8)
        for wrap in (False, True):
            data = qs.encode_qoi2avg(img, par=2.0, dpi=72.0)
            if wrap:
                data = qs.lz4_wrap(data, oracle)
            px, d, t = oracle.qoix_decode(data, 0)
            assert np.array_equal(px, img) and t == (9 if c == 3 else 12)
            assert (d.width, d.height, d.pitchBytes, d.pixelAspectRatio, d.resolutionY) == (w, h, w * c, 2.0, 72.0)
        r = oracle.qoix_decode(qs.fuzz_qoi2avg(w, h, c, 5), 0)
        assert r is not None and r[0].shape == (h, w, c)


@pytest.mark.parametrize("c", [1, 2])
def test_qoiplane_literal_roundtrip_and_fuzz(oracle, c):
    import qoixsynth as qs
    for (h, w) in [(1, 1), (9, 13), (40, 31)]:
        img = _rand_img(h, w, c, h * w + c).astype(np.uint8)
        px, d, t = oracle.qoix_decode(qs.encode_qoiplane(img), 0)
        assert np.array_equal(px, img) and t == (0 if c == 1 else 3)
        r = oracle.qoix_decode(qs.fuzz_qoiplane(w, h, c, 6), 0)
        assert r is not None and r[0].shape == (h, w, c)


@pytest.mark.parametrize("c", [1, 2, 3, 4])
@pytest.mark.parametrize("version", [1, 2])
def test_qoi10b_literal_roundtrip_and_fuzz(oracle, c, version):
    import qoixsynth as qs
    if version == 2 and c <= 2:
        pytest.skip("10-bit L/LA version 2 streams are QOI-Plane10 (plugins/qoix.d:438-447)")
    for (h, w) in [(1, 1), (9, 13), (40, 31)]:
        v = _rand_img(h, w, c, h * w + c, 1024)
        exp = ((v << 6) | (v >> 4)).astype(np.uint16)
        px, d, t = oracle.qoix_decode(qs.encode_qoi10b(v, version), 0)
        assert np.array_equal(px, exp) and t == {1: 1, 2: 4, 3: 10, 4: 13}[c]
        r = oracle.qoix_decode(qs.fuzz_qoi10b(w, h, c, 7, version), 0)
        assert r is not None and r[0].shape == (h, w, c)


def test_sub_codec_rejects(oracle):
    import qoixsynth as qs
    img = _rand_img(4, 4, 3, 1).astype(np.uint8)
    good = qs.encode_qoi2avg(img)
    assert oracle.qoix_decode(good, 0) is not None
    for off, val in ((12, 2), (15, 3), (0, 0x70), (16, 2)):
        bad = bytearray(good); bad[off] = val
        assert oracle.qoix_decode(bytes(bad), 0) is None
    assert oracle.qoix_decode(good[:27], 0) is None


@pytest.mark.parametrize("c", [1, 2, 3, 4])
def test_8bit_sub_encoders_round_trip(oracle, c):
    """or_qoiplane_encode (qoiplane.d:109-375, c = 1|2) and or_qoix_encode (qoi2avg.d:376-617, c = 3|4): the reference's
    round-trip property (image.d:2112-2183) through the restated decoders, on images that reach every opcode, the run
    limits and -- for QOI2AVG -- the eviction of its 64-entry FIFO index; header fields survive; the refusals of the
    cited checks."""
    rng = np.random.default_rng(40 + c)
    if c <= 2:
        from test_qoix_encode_emulated import plane8_images
        imgs, enc = plane8_images(c, rng), oracle.qoiplane_encode
    else:
        from test_qoi2avg_encode_emulated import qoi2avg_images
        imgs, enc = qoi2avg_images(c, rng), oracle.qoi2avg_encode
    for img in imgs:
        s = enc(img, colorspace=1, par=1.25, dpi=300.0)
        assert s is not None and s[:4] == b"qoix" and s[12] == 1 and s[13] == c and s[14] == 8 and s[15] == 1 and s[16] == 0
        px, desc, _ = oracle.qoix_decode(s, 0)
        assert np.array_equal(px, img) and desc.pixelAspectRatio == 1.25 and desc.resolutionY == 300.0
    if c >= 3:
        assert oracle.qoi2avg_encode(imgs[4], colorspace=3) is None           # desc.colorspace > 2 (qoi2avg.d:392)
        assert oracle.qoiplane_encode(imgs[4]) is None                        # channels 3 / 4 are not QOI-Plane's (qoiplane.d:111)
    else:
        assert oracle.qoi2avg_encode(imgs[4]) is None                         # channels 1 / 2 are not QOI2AVG's (qoi2avg.d:391)
