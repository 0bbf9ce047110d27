"""gb200_decode_batch_host (host files in, host pixels out, transfers pipelined over sub-batches) against the oracle:
every image of a mixed batch, including a corrupt one in the middle, for the three batched formats."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codecs(gb):
    from gamut_b200 import codecs
    return codecs


def _run(codecs, fmt, files, arg, stride, sub):
    from gamut_b200 import _lib
    L = _lib.lib()
    host = L.gb200_host_alloc(stride * len(files))
    assert host
    try:
        descs = codecs.decode_batch_host(fmt, files, arg, -1 if fmt == 1 else 0, host, stride, sub)
        out = []
        for i, d in enumerate(descs):
            if not d.status:
                out.append(None)
                continue
            assert d.pixels == host + i * stride
            n = d.pitch * d.height
            out.append((np.ctypeslib.as_array(C.cast(d.pixels, C.POINTER(C.c_uint8)), shape=(n,)).copy(), d))
        return out
    finally:
        L.gb200_host_free(host)


@pytest.mark.parametrize("sub", [0, 1, 3])
def test_jpeg_batch_host(codecs, oracle, sub):
    import jpegutil
    files = [jpegutil.encode(jpegutil.photo(40 + 8 * i, 64 + 16 * i, 3, i), 90, (2, 0, 1)[i % 3]) for i in range(7)]
    files.insert(3, b"\xff\xd8\xff garbage")
    res = _run(codecs, 0, files, -1, 256 * 256 * 3, sub)
    for i, f in enumerate(files):
        exp = oracle.jpeg_load(f, -1)
        if exp is None:
            assert res[i] is None
            continue
        px, d = res[i]
        assert (d.width, d.height, d.channels) == (exp[0].shape[1], exp[0].shape[0], exp[0].shape[2])
        assert np.array_equal(px, exp[0].reshape(-1)), i


def test_png_batch_host(codecs, oracle):
    from pngwriter import write_png
    rng = np.random.default_rng(3)
    files = [write_png(rng.integers(0, 256, (20 + i, 33 + i, 4)).astype(np.uint8), 6, 8, filters="adaptive") for i in range(5)]
    files.append(write_png(rng.integers(0, 65536, (9, 14, 2)), 4, 16, filters=4))
    files.insert(2, files[0][:40])
    res = _run(codecs, 1, files, 0, 64 * 64 * 8, 2)
    for i, f in enumerate(files):
        exp, info = oracle.png_load(f, 0, 1 if oracle.png_is16(f) else 0)
        if exp is None:
            assert res[i] is None
            continue
        px, d = res[i]
        assert np.array_equal(px, exp.reshape(-1).view(np.uint8)), i


def test_qoix_batch_host(codecs, oracle):
    from qoixutil import depth_map_la
    files = [oracle.qoix_encode(depth_map_la(24 + i, 40 + 3 * i, i, 2), 10, force_lz4=bool(i & 1)) for i in range(6)]
    files.insert(1, b"qoix" + b"\0" * 40)
    res = _run(codecs, 3, files, 0, 64 * 64 * 4, 4)
    for i, f in enumerate(files):
        exp = oracle.qoix_decode(f, 0)
        if exp is None:
            assert res[i] is None
            continue
        px, d = res[i]
        assert np.array_equal(px, exp[0].reshape(-1).view(np.uint8)), i


def test_stride_too_small_fails_that_image_only(codecs):
    import jpegutil
    files = [jpegutil.encode(jpegutil.photo(16, 16, 3, 1), 90, 0), jpegutil.encode(jpegutil.photo(64, 64, 3, 2), 90, 0)]
    res = _run(codecs, 0, files, -1, 16 * 16 * 3, 0)
    assert res[0] is not None and res[1] is None


def test_png_batch_host_sliced(codecs, oracle):
    """100 files with the default sub-batch: the PNG pipeline runs as slices side by side from several host threads
    (batch_host.cu); every image, including a corrupt one in each slice, must come out as from a single call."""
    from pngwriter import write_png
    rng = np.random.default_rng(11)
    base = [write_png(rng.integers(0, 256, (12 + i, 17 + 2 * i, 4)).astype(np.uint8), 6, 8, filters="adaptive") for i in range(10)]
    files = [base[i % 10] for i in range(100)]
    for k in (5, 50, 95):
        files[k] = files[k][:60]
    res = _run(codecs, 1, files, 0, 64 * 64 * 4, 0)
    exp10 = [oracle.png_load(f, 0, 0)[0] for f in base]
    for i in range(100):
        if i in (5, 50, 95):
            assert res[i] is None
            continue
        px, d = res[i]
        assert np.array_equal(px, exp10[i % 10].reshape(-1).view(np.uint8)), i
