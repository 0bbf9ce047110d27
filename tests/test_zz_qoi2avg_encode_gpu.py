"""GPU parity of the QOI2AVG encoder: gb200_qoix_encode on rgb8 / rgba8 images must produce, byte for byte,
the stream of the reference's qoix_encode (codecs/qoi2avg.d:376-617, restated in oracle/qoix_sub_oracle.c), and both
decoders must read it back to the original pixels (the round trip of the reference's own test, image.d:2112-2183).
(Developed under the CPU emulation, tests/test_qoi2avg_encode_emulated.py; first GPU run: profiles/r2_qoi2avg_pytest.txt.)"""
import numpy as np
import pytest

from qoixutil import qoi_test_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codecs(gb):
    from gamut_b200 import codecs
    return codecs


def check(codecs, oracle, img, **kw):
    exp = oracle.qoi2avg_encode(img, **kw)
    got = codecs.qoix_encode(img, **kw)
    assert exp is not None and got is not None
    assert len(got) == len(exp)
    assert got == exp
    dec = codecs.qoix_decode(got)
    assert dec is not None and np.array_equal(dec[0], img)
    assert np.array_equal(oracle.qoix_decode(got, 0)[0], img)
    return got


@pytest.mark.parametrize("c", [3, 4])
def test_every_opcode_class_and_the_fifo(codecs, oracle, c):
    from test_qoi2avg_encode_emulated import qoi2avg_images
    for img in qoi2avg_images(c, np.random.default_rng(30 + c)):
        check(codecs, oracle, img, par=1.5, dpi=96.0, colorspace=1)
    check(codecs, oracle, qoi_test_image(257, 1024, c, 5))
    check(codecs, oracle, qoi_test_image(1080, 1920, c, 6))


def test_pitch_mixed_batch_and_rejects(codecs, oracle):
    import ctypes as C
    import torch
    from gamut_b200 import codecs as cd
    from qoixutil import depth_map_la
    rng = np.random.default_rng(3)
    img = qoi_test_image(20, 30, 4, 1)
    wide = rng.integers(0, 256, (20, 37, 4)).astype(np.uint8)
    wide[:, :30] = img                                             # row padding must not be read as pixels
    n = C.c_int(0)
    p = cd._L().gb200_qoix_encode(wide.ctypes.data, C.byref(cd.QoixDesc(30, 20, 148, 4, 8, 0, 0, -1.0, -1.0)), C.byref(n))
    assert p and cd._take_host(p, n.value).tobytes() == oracle.qoi2avg_encode(img)
    for bad in (cd.QoixDesc(30, 20, 148, 4, 10, 0, 0, -1, -1), cd.QoixDesc(30, 20, 148, 4, 8, 3, 0, -1, -1), cd.QoixDesc(30, 20, 148, 4, 8, 0, 1, -1, -1),
                cd.QoixDesc(30, 20, 119, 4, 8, 0, 0, -1, -1), cd.QoixDesc(0, 20, 148, 4, 8, 0, 0, -1, -1), cd.QoixDesc(30, 20, 148, 5, 8, 0, 0, -1, -1)):
        assert not cd._L().gb200_qoix_encode(wide.ctypes.data, C.byref(bad), C.byref(n))
    # one device batch holding all three sub-codecs in any order
    imgs = [qoi_test_image(300, 500, 3, 6), depth_map_la(200, 333, 7, 2), qoi_test_image(512, 512, 4, 8), (depth_map_la(64, 64, 8, 1) >> 8).astype(np.uint8)]
    exp = [oracle.qoi2avg_encode(imgs[0]), oracle.qoiplane10_encode(imgs[1]), oracle.qoi2avg_encode(imgs[2]), oracle.qoiplane_encode(imgs[3])]
    dev = [torch.from_numpy(i.view(np.int16) if i.itemsize == 2 else i).cuda() for i in imgs]
    outs = [torch.empty(i.shape[0] * i.shape[1] * 5 + 256, dtype=torch.uint8, device="cuda") for i in imgs]
    lens = codecs.qoix_encode_batch_device([t.data_ptr() for t in dev], [i.shape for i in imgs], [o.data_ptr() for o in outs], bitdepths=[8, 10, 8, 8])
    torch.cuda.synchronize()
    for o, k, e in zip(outs, lens, exp):
        assert k == len(e) and o[:k].cpu().numpy().tobytes() == e


def test_image_save_qoix_rgb(codecs, oracle):
    """Image.saveToMemory(QOIX) of an rgb8 / rgba8 image (saveQOIX -> qoix_encode, plugins/qoix.d:185-198)."""
    from gamut_b200.image import Image
    from gamut_b200.types import ImageFormat, PixelType
    for c in (3, 4):
        img = qoi_test_image(37, 61, c, 9)
        src = oracle.qoi2avg_encode(img, par=2.0, dpi=72.0)
        im = Image()
        assert im.loadFromMemory(src, 0) and im.type() == (PixelType.rgb8 if c == 3 else PixelType.rgba8)
        assert im.saveToMemory(ImageFormat.QOIX) == src
