"""GPU parity of the QOI-Plane10 encoder: gb200_qoix_encode must produce, byte for byte, the stream of the
reference's qoiplane10_encode (codecs/qoiplane10.d:99-314, restated line by line in oracle/qoix_oracle.c), and both
decoders must read it back to the original pixels (the round trip the reference's own test does, image.d:2112-2183)."""
import numpy as np
import pytest

from qoixutil import depth_map_la

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codecs(gb):
    from gamut_b200 import codecs
    return codecs


def expand(v):
    return ((v << 6) | (v >> 4)).astype(np.uint16)


def check(codecs, oracle, img, **kw):
    exp = oracle.qoiplane10_encode(img, **kw)
    got = codecs.qoix_encode(img, **kw)
    assert exp is not None and got is not None
    assert len(got) == len(exp)
    assert got == exp
    dec = codecs.qoix_decode(got)
    assert dec is not None and np.array_equal(dec[0], img)
    assert np.array_equal(oracle.qoix_decode(got, 0)[0], img)
    return got


@pytest.mark.parametrize("c", [1, 2])
def test_depth_maps(codecs, oracle, c):
    for (h, w) in [(1, 1), (1, 2), (2, 1), (3, 5), (33, 47), (2, 300), (64, 64), (200, 333), (257, 1024)]:
        check(codecs, oracle, depth_map_la(h, w, 3 + h, c), par=1.5, dpi=96.0)


@pytest.mark.parametrize("c", [1, 2])
def test_every_opcode_class(codecs, oracle, c):
    rng = np.random.default_rng(c)
    h, w = 70, 91
    # noise: DIFF4 / LA everywhere
    check(codecs, oracle, expand(rng.integers(0, 1024, (h, w, c))))
    # small steps: DIFF1 / DIFF2 / DIFF3 and short alpha differences
    base = np.cumsum(rng.integers(-40, 41, (h, w, c)), axis=1) % 1024
    check(codecs, oracle, expand(base))
    # runs of every length around the 7 / 256 / row boundaries, including runs of one with a tiny residual
    v = np.zeros((h * w, c), np.int64)
    pos = 0
    for n in [1, 1, 2, 6, 7, 8, 9, 255, 256, 257, 300, 513, 1, 3, 700]:
        if pos >= h * w:
            break
        v[pos:pos + n] = rng.integers(0, 1024, c)
        pos += n
    v[pos:] = rng.integers(0, 1024, (h * w - pos, c))
    check(codecs, oracle, expand(v.reshape(h, w, c)))
    # one flat image: a single sequence cut every 256 pixels, across rows and tiles
    check(codecs, oracle, expand(np.full((40, 130, c), 517)))
    check(codecs, oracle, expand(np.zeros((40, 130, c), np.int64) + np.array([0, 1023])[:c]))   # equal to the initial predictor


def test_pitch_colorspace_and_rejects(codecs, oracle):
    img = depth_map_la(20, 30, 1, 2)
    wide = np.zeros((20, 40, 2), np.uint16)
    wide[:, :30] = img
    wide[:, 30:] = 0xABCD                                         # row padding must not be read as pixels
    exp = oracle.qoiplane10_encode(img, colorspace=1)
    from gamut_b200 import codecs as cd
    import ctypes as C
    d = cd.QoixDesc(30, 20, 40 * 4, 2, 10, 1, 0, -1.0, -1.0)
    n = C.c_int(0)
    p = cd._L().gb200_qoix_encode(wide.ctypes.data, C.byref(d), C.byref(n))
    assert p and cd._take_host(p, n.value).tobytes() == exp
    for bad in (cd.QoixDesc(30, 20, 120, 3, 10, 0, 0, -1, -1), cd.QoixDesc(30, 20, 120, 2, 9, 0, 0, -1, -1), cd.QoixDesc(30, 20, 120, 4, 10, 0, 0, -1, -1),
                cd.QoixDesc(0, 20, 120, 2, 10, 0, 0, -1, -1), cd.QoixDesc(30, 20, 120, 2, 10, 0, 1, -1, -1),
                cd.QoixDesc(30, 20, 60, 2, 10, 0, 0, -1, -1)):
        assert not cd._L().gb200_qoix_encode(wide.ctypes.data, C.byref(bad), C.byref(n))


def test_config5_shape_2048_and_batch(codecs, oracle, gb):
    """BASELINE config 5 shape: 2048x2048 10-bit LA, single image through the host API and a device-resident batch."""
    import torch
    imgs = [depth_map_la(2048, 2048, 5, 2), depth_map_la(300, 500, 6, 1), depth_map_la(2048, 2048, 7, 2)]
    exp = [oracle.qoiplane10_encode(i) for i in imgs]
    assert codecs.qoix_encode(imgs[0]) == exp[0]
    dev = [torch.from_numpy(i.view(np.int16)).cuda() for i in imgs]
    outs = [torch.empty(codecs.qoix_encode_bound(i.shape[1], i.shape[0], i.shape[2]) + 16, dtype=torch.uint8, device="cuda") for i in imgs]
    lens = codecs.qoix_encode_batch_device([t.data_ptr() for t in dev], [i.shape for i in imgs], [o.data_ptr() for o in outs])
    torch.cuda.synchronize()
    for o, n, e in zip(outs, lens, exp):
        assert n == len(e) and o[:n].cpu().numpy().tobytes() == e


def test_image_save_qoix(codecs, oracle):
    """Image.saveToMemory(QOIX) of a loaded 10-bit image (saveQOIX, plugins/qoix.d:156-241): the stream the reference's
    encoder writes for the image's pixels, metadata included, and it loads back to the same image."""
    from gamut_b200.image import Image
    from gamut_b200.types import ImageFormat, PixelType
    for c in (1, 2):
        img = depth_map_la(37, 61, 9, c)
        src = oracle.qoiplane10_encode(img, par=2.0, dpi=72.0)
        im = Image()
        assert im.loadFromMemory(src, 0) and im.type() == (PixelType.l16 if c == 1 else PixelType.la16)
        out = im.saveToMemory(ImageFormat.QOIX)
        assert out == src
        assert im.saveToMemory(ImageFormat.PNG) is None


# ---- QOI-Plane (8-bit L / LA): qoiplane_encode, codecs/qoiplane.d:109-375 --------------------------------------------
def check8(codecs, oracle, img, **kw):
    exp = oracle.qoiplane_encode(img, **kw)
    got = codecs.qoix_encode(img, **kw)
    assert exp is not None and got is not None
    assert len(got) == len(exp)
    assert got == exp
    dec = codecs.qoix_decode(got)
    assert dec is not None and np.array_equal(dec[0], img)
    assert np.array_equal(oracle.qoix_decode(got, 0)[0], img)
    return got


@pytest.mark.parametrize("c", [1, 2])
def test_qoiplane_8bit(codecs, oracle, c):
    from test_qoix_encode_emulated import plane8_images
    for img in plane8_images(c, np.random.default_rng(10 + c)):
        check8(codecs, oracle, img, par=1.5, dpi=96.0, colorspace=1)
    check8(codecs, oracle, (depth_map_la(257, 1024, 5, c) >> 8).astype(np.uint8))
    check8(codecs, oracle, (depth_map_la(1080, 1920, 6, c) >> 8).astype(np.uint8))


def test_qoiplane_8bit_pitch_and_mixed_batch(codecs, oracle):
    import ctypes as C
    import torch
    from gamut_b200 import codecs as cd
    rng = np.random.default_rng(3)
    a8 = (depth_map_la(20, 30, 1, 2) >> 8).astype(np.uint8)
    wide = rng.integers(0, 256, (20, 37, 2)).astype(np.uint8)
    wide[:, :30] = a8                                              # row padding must not be read as pixels
    d = cd.QoixDesc(30, 20, 74, 2, 8, 0, 0, -1.0, -1.0)
    n = C.c_int(0)
    p = cd._L().gb200_qoix_encode(wide.ctypes.data, C.byref(d), C.byref(n))
    assert p and cd._take_host(p, n.value).tobytes() == oracle.qoiplane_encode(a8)
    assert not cd._L().gb200_qoix_encode(wide.ctypes.data, C.byref(cd.QoixDesc(30, 20, 59, 2, 8, 0, 0, -1, -1)), C.byref(n))
    # one device batch holding 10-bit and 8-bit images in any order
    imgs = [(depth_map_la(300, 500, 6, 1) >> 8).astype(np.uint8), depth_map_la(200, 333, 7, 2), (depth_map_la(512, 512, 8, 2) >> 8).astype(np.uint8)]
    exp = [oracle.qoiplane_encode(imgs[0]), oracle.qoiplane10_encode(imgs[1]), oracle.qoiplane_encode(imgs[2])]
    dev = [torch.from_numpy(i.view(np.int16) if i.itemsize == 2 else i).cuda() for i in imgs]
    outs = [torch.empty(codecs.qoix_encode_bound(i.shape[1], i.shape[0], i.shape[2]) + 16, dtype=torch.uint8, device="cuda") for i in imgs]
    lens = codecs.qoix_encode_batch_device([t.data_ptr() for t in dev], [i.shape for i in imgs], [o.data_ptr() for o in outs], bitdepths=[8, 10, 8])
    torch.cuda.synchronize()
    for o, k, e in zip(outs, lens, exp):
        assert k == len(e) and o[:k].cpu().numpy().tobytes() == e


def test_image_save_qoix_8bit(codecs, oracle):
    """Image.saveToMemory(QOIX) of an 8-bit greyscale image (saveQOIX -> qoiplane_encode, plugins/qoix.d:172-184)."""
    from gamut_b200.image import Image
    from gamut_b200.types import ImageFormat, PixelType
    for c in (1, 2):
        img = (depth_map_la(37, 61, 9, c) >> 8).astype(np.uint8)
        src = oracle.qoiplane_encode(img, par=2.0, dpi=72.0)
        im = Image()
        assert im.loadFromMemory(src, 0) and im.type() == (PixelType.l8 if c == 1 else PixelType.la8)
        assert im.saveToMemory(ImageFormat.QOIX) == src
