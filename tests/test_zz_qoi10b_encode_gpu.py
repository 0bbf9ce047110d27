"""GPU parity of the QOI-10b encoder: gb200_qoix_encode on rgb16 / rgba16 images must produce, byte for byte,
the stream of the reference's qoi10b_encode (codecs/qoi10b.d:136-500, restated in oracle/qoix_sub_oracle.c), and both
decoders must read it back to the original pixels.

These kernels were written after the round's GPU budget was spent: they are byte-exact under the CPU emulation
(tests/test_qoi10b_encode_emulated.py, also with AddressSanitizer) like the five kernel sets before them, all of which
then passed on their first GPU run, but they have not run on a GPU yet. Until they have, the tests are marked
xfail(strict=False): a pass shows as XPASS, a failure cannot hide the rest of the suite."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="first GPU run pending (developed under the CPU emulation after the GPU budget ended)")]


@pytest.fixture(scope="module")
def codecs(gb):
    from gamut_b200 import codecs
    return codecs


def check(codecs, oracle, img, **kw):
    exp = oracle.qoi10b_encode(img, **kw)
    got = codecs.qoix_encode(img, **kw)
    assert exp is not None and got is not None
    assert len(got) == len(exp)
    assert got == exp
    dec = codecs.qoix_decode(got)
    assert dec is not None and np.array_equal(dec[0], img)
    assert np.array_equal(oracle.qoix_decode(got, 0)[0], img)
    return got


@pytest.mark.parametrize("c", [3, 4])
def test_every_opcode_class(codecs, oracle, c):
    from test_qoi10b_encode_emulated import qoi10b_images
    from test_qoix_encode_emulated import expand
    rng = np.random.default_rng(50 + c)
    for img in qoi10b_images(c, rng):
        check(codecs, oracle, img, par=1.5, dpi=96.0, colorspace=1)
    check(codecs, oracle, expand(np.cumsum(rng.integers(-6, 7, (1080, 1920, c)), axis=1) % 1024))


def test_pitch_batch_and_image(codecs, oracle):
    import ctypes as C
    import torch
    from gamut_b200 import codecs as cd
    from gamut_b200.image import Image
    from gamut_b200.types import ImageFormat, PixelType
    from test_qoix_encode_emulated import expand
    rng = np.random.default_rng(3)
    img = expand(rng.integers(0, 1024, (20, 30, 4)))
    wide = rng.integers(0, 65536, (20, 37, 4)).astype(np.uint16)
    wide[:, :30] = img                                             # row padding must not be read as pixels
    n = C.c_int(0)
    p = cd._L().gb200_qoix_encode(wide.ctypes.data, C.byref(cd.QoixDesc(30, 20, 296, 4, 10, 0, 0, -1.0, -1.0)), C.byref(n))
    assert p and cd._take_host(p, n.value).tobytes() == oracle.qoi10b_encode(img)
    imgs = [expand(rng.integers(0, 1024, (300, 500, 3))), expand(np.cumsum(rng.integers(-3, 4, (512, 512, 4)), axis=1) % 1024)]
    exp = [oracle.qoi10b_encode(i) for i in imgs]
    dev = [torch.from_numpy(i.view(np.int16)).cuda() for i in imgs]
    outs = [torch.empty(i.shape[0] * i.shape[1] * 7 + 256, dtype=torch.uint8, device="cuda") for i in imgs]
    lens = codecs.qoix_encode_batch_device([t.data_ptr() for t in dev], [i.shape for i in imgs], [o.data_ptr() for o in outs], bitdepths=[10, 10])
    torch.cuda.synchronize()
    for o, k, e in zip(outs, lens, exp):
        assert k == len(e) and o[:k].cpu().numpy().tobytes() == e
    for c in (3, 4):                                               # Image.saveToMemory(QOIX) of an rgb16 / rgba16 image
        im16 = expand(np.cumsum(rng.integers(-9, 10, (37, 61, c)), axis=1) % 1024)
        src = oracle.qoi10b_encode(im16, par=2.0, dpi=72.0)
        im = Image()
        assert im.loadFromMemory(src, 0) and im.type() == (PixelType.rgb16 if c == 3 else PixelType.rgba16)
        assert im.saveToMemory(ImageFormat.QOIX) == src
