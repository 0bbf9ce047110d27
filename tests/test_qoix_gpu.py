"""GPU parity: CUDA QOI / LZ4 / QOI-Plane10 (through the C ABI) vs the CPU oracle, bit-exact.
Reference: codecs/qoi.d, codecs/lz4.d, codecs/qoiplane10.d, plugins/qoix.d."""
import numpy as np
import pytest

from qoixutil import depth_map_la, qoi_bytes, qoi_test_image
import qoixsynth  # noqa: F401  (tests/ is on sys.path)

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


# ---- remaining QOIX sub-codecs: QOI2AVG (8-bit RGB/RGBA), QOI-Plane (8-bit L/LA), QOI-10b ----
SIZES = [(1, 1), (3, 2), (9, 13), (40, 31), (64, 128)]


@pytest.mark.parametrize("c", [3, 4])
def test_qoi2avg(codecs, oracle, c):
    import qoixsynth as qs
    rng = np.random.default_rng(c)
    for (h, w) in SIZES:
        img = rng.integers(0, 256, (h, w, c)).astype(np.uint8)
        got = check_qoix(codecs, oracle, qs.encode_qoi2avg(img, par=1.25, dpi=300.0))
        assert np.array_equal(got[0], img)
        for seed in range(4):
            s = qs.fuzz_qoi2avg(w, h, c, 100 * seed + h)
            assert check_qoix(codecs, oracle, s) is not None
            assert check_qoix(codecs, oracle, qs.lz4_wrap(s, oracle)) is not None
    check_qoix(codecs, oracle, s[:40] + b"\xff" * 4)                # truncated: stream ends early
    check_qoix(codecs, oracle, s[:30] + b"\xff\xff" + s[32:])      # END opcode in the middle of a row


@pytest.mark.parametrize("c", [1, 2])
def test_qoiplane8(codecs, oracle, c):
    import qoixsynth as qs
    rng = np.random.default_rng(c)
    for (h, w) in SIZES:
        img = rng.integers(0, 256, (h, w, c)).astype(np.uint8)
        got = check_qoix(codecs, oracle, qs.encode_qoiplane(img))
        assert np.array_equal(got[0], img)
        for seed in range(4):
            s = qs.fuzz_qoiplane(w, h, c, 100 * seed + h)
            assert check_qoix(codecs, oracle, s) is not None
            assert check_qoix(codecs, oracle, qs.lz4_wrap(s, oracle)) is not None
    check_qoix(codecs, oracle, s[:40] + b"\xff" * 4)


@pytest.mark.parametrize("c,version", [(1, 1), (2, 1), (3, 1), (4, 1), (3, 2), (4, 2), (1, 0)])
def test_qoi10b(codecs, oracle, c, version):
    import qoixsynth as qs
    rng = np.random.default_rng(c + 10 * version)
    for (h, w) in SIZES:
        v = rng.integers(0, 1024, (h, w, c))
        got = check_qoix(codecs, oracle, qs.encode_qoi10b(v, version))
        assert np.array_equal(got[0], ((v << 6) | (v >> 4)).astype(np.uint16))
        for seed in range(4):
            s = qs.fuzz_qoi10b(w, h, c, 100 * seed + h, version)
            assert check_qoix(codecs, oracle, s) is not None
            assert check_qoix(codecs, oracle, qs.lz4_wrap(s, oracle)) is not None
    check_qoix(codecs, oracle, s[:60] + b"\xff" * 5)               # early END: remaining rows stay zero


def test_sub_codec_batch_and_rejects(codecs, oracle):
    import qoixsynth as qs
    files = [qs.fuzz_qoi2avg(33, 17, 4, 1), qs.fuzz_qoiplane(20, 50, 2, 2), qs.fuzz_qoi10b(31, 9, 3, 3, 2),
             oracle.qoix_encode(depth_map_la(40, 50, 1, 2), 10, force_lz4=True), qs.fuzz_qoi10b(8, 8, 1, 4, 1)]
    bad = bytearray(files[0]); bad[12] = 2
    files.append(bytes(bad))                                          # version 2 of an 8-bit stream: rejected
    bad = bytearray(files[1]); bad[15] = 2
    files.append(bytes(bad))                                          # premultiplied QOI-Plane: rejected
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


def test_lz4_long_blocks_parallel_walk(codecs, oracle):
    """Blocks long enough (>= 64 KB of LZ4 bytes) for the sub-chunk-parallel walk (lz4_spec/merge/scan/pwrite), plus
    damaged copies that must fail -- or decode identically -- in both implementations."""
    rng = np.random.default_rng(5)
    for (h, w, seed) in [(768, 1024, 1), (600, 900, 2)]:
        img = depth_map_la(h, w, seed, 2)
        data = oracle.qoix_encode(img, 10, force_lz4=True)
        assert len(data) > 4 * 16384 + 29
        got = check_qoix(codecs, oracle, data)
        assert np.array_equal(got[0], img)
        for _ in range(6):
            b = bytearray(data)
            i = int(rng.integers(40, len(b) - 8))
            b[i] ^= 1 << int(rng.integers(0, 8))
            check_qoix(codecs, oracle, bytes(b))
        check_qoix(codecs, oracle, data[:len(data) // 2])
        check_qoix(codecs, oracle, data + b"\0" * 37)          # trailing bytes after the final sequence


@pytest.mark.parametrize("mode", ["1", "2"])
def test_plane10_boundary_repair_paths(codecs, oracle, monkeypatch, mode):
    """The entry of every sync CTA's first own chunk is declared wrong (test hook): mode 1 sends every boundary
    through p10_repair_kernel's forward walk, mode 2 leaves that to the host loop that runs when the device-side
    check still counts disagreeing boundaries. Both must end at the serial parse."""
    monkeypatch.setenv("GB200_P10_FORCE_REPAIR", mode)
    for (h, w, c) in [(700, 901, 2), (512, 768, 1)]:
        img = depth_map_la(h, w, 7, c)
        for force in (False, True):
            got = check_qoix(codecs, oracle, oracle.qoix_encode(img, 10, force_lz4=force))
            assert np.array_equal(got[0], img)
