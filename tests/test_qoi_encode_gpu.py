"""GPU parity of the QOI encoder (the save side of configs[0]): gb200_qoi_encode must produce, byte for
byte, the stream of the reference's qoi_encode (codecs/qoi.d:295-426, restated in oracle/qoix_oracle.c and pinned to PIL's
independent writer in tests/test_oracle_qoix.py), and both decoders must read it back to the original pixels (the round
trip the reference's own test does, image.d:2112-2183)."""
import numpy as np
import pytest

from qoixutil import qoi_test_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codecs(gb):
    from gamut_b200 import codecs
    return codecs


def check(codecs, oracle, img, **kw):
    exp = oracle.qoi_encode(img, **kw)
    got = codecs.qoi_encode(img, **kw)
    assert exp is not None and got is not None
    assert len(got) == len(exp)
    assert got == exp
    dec = codecs.qoi_decode(got, 0)
    assert dec is not None and np.array_equal(dec[0], img)
    assert np.array_equal(oracle.qoi_decode(got, 0)[0], img)
    return got


@pytest.mark.parametrize("c", [3, 4])
def test_test_images(codecs, oracle, c):
    for (h, w) in [(1, 1), (1, 2), (2, 1), (3, 5), (33, 47), (2, 300), (64, 64), (200, 333), (257, 1024)]:
        check(codecs, oracle, qoi_test_image(h, w, c, 3 + h), colorspace=h & 1)


def test_3x1_kat(codecs, oracle):
    # image.d:2112-2183: 3x1 rgb8 [255,0,0, 15,64,255, 0,255,255] must survive encode -> decode
    check(codecs, oracle, np.array([[[255, 0, 0], [15, 64, 255], [0, 255, 255]]], np.uint8))


@pytest.mark.parametrize("c", [3, 4])
def test_every_opcode_class(codecs, oracle, c):
    rng = np.random.default_rng(c)
    h, w = 70, 91
    check(codecs, oracle, rng.integers(0, 256, (h, w, c)).astype(np.uint8))                      # noise: RGB / RGBA
    base = (np.cumsum(rng.integers(-3, 4, (h, w, c)), axis=1) % 256).astype(np.uint8)            # small steps: DIFF / LUMA
    check(codecs, oracle, base)
    base = (np.cumsum(rng.integers(-20, 21, (h, w, c)), axis=1) % 256).astype(np.uint8)          # LUMA / RGB
    check(codecs, oracle, base)
    # runs of every length around the 62 / row / tile boundaries
    v = np.zeros((h * w, c), np.uint8)
    pos = 0
    for n in [1, 1, 2, 61, 62, 63, 124, 125, 300, 1, 3, 1024, 1100]:
        v[pos:pos + n] = rng.integers(0, 256, c)
        pos += n
    v[pos:] = rng.integers(0, 256, (h * w - pos, c))
    check(codecs, oracle, v.reshape(h, w, c))
    # one flat image: a single sequence cut every 62 pixels, across rows and tiles; and one equal to the initial px_prev
    check(codecs, oracle, np.full((40, 130, c), 200, np.uint8))
    first = np.zeros((40, 130, c), np.uint8)
    first[..., 3:] = 255
    check(codecs, oracle, first)
    check(codecs, oracle, np.zeros((40, 130, c), np.uint8))                                      # rgba8: the zeroed index hits
    # index hits inside a tile and across tiles: few colours, several colours per bucket, the all-zero pixel
    pal = rng.integers(0, 256, (5, c)).astype(np.uint8)
    check(codecs, oracle, pal[rng.integers(0, 5, 90 * 90)].reshape(90, 90, c))
    pal = rng.integers(0, 256, (90, c)).astype(np.uint8)
    check(codecs, oracle, pal[rng.integers(0, 90, 90 * 90)].reshape(90, 90, c))
    z = rng.integers(0, 256, (40, 40, c)).astype(np.uint8)
    z[5::7] = 0
    check(codecs, oracle, z)


def test_pitch_flip_alignment_and_rejects(codecs, oracle):
    rng = np.random.default_rng(9)
    img = qoi_test_image(37, 45, 4, 4)
    exp = oracle.qoi_encode(img)
    wide = rng.integers(0, 256, (37, 60, 4)).astype(np.uint8)
    wide[:, :45] = img                                             # row padding must not be read as pixels
    assert codecs.qoi_encode(wide, pitch=240, shape=(37, 45, 4)) == exp
    flipped = np.ascontiguousarray(wide[::-1])                     # saveQOI passes image._pitch, negative when flipped
    assert codecs.qoi_encode(flipped, pitch=-240, first_scanline=36 * 240, shape=(37, 45, 4)) == exp
    buf = np.zeros(img.size + 8, np.uint8)                         # rgba8 at an odd address: byte loads
    o = (-buf.ctypes.data) % 4 + 1
    buf[o:o + img.size] = img.reshape(-1)
    assert codecs.qoi_encode(buf, first_scanline=o, shape=(37, 45, 4)) == exp
    img3 = qoi_test_image(20, 31, 3, 6)
    wide3 = rng.integers(0, 256, (20, 98), dtype=np.uint8)
    wide3[:, :93] = img3.reshape(20, 93)
    assert codecs.qoi_encode(wide3, pitch=98, shape=(20, 31, 3)) == oracle.qoi_encode(img3)
    # qoi_encode's refusals (qoi.d:303-311): channels, colorspace, empty; and a pitch smaller than a scanline
    ok = qoi_test_image(4, 4, 4, 1)
    assert codecs.qoi_encode(ok, shape=(4, 4, 2), pitch=8) is None
    assert codecs.qoi_encode(ok, shape=(4, 4, 5), pitch=20) is None
    assert codecs.qoi_encode(ok, shape=(0, 4, 4)) is None and codecs.qoi_encode(ok, shape=(4, 0, 4)) is None
    assert codecs.qoi_encode(ok, colorspace=2) is None
    assert codecs.qoi_encode(ok, pitch=8) is None
    assert codecs.qoi_encode(ok) == oracle.qoi_encode(ok)


def test_config1_shape_and_batch_device(codecs, oracle):
    """The 512x512 RGBA8 image of BASELINE configs[0] and larger images, device-resident, in one batch; a refused image
    in the middle of the batch gets length 0 and does not disturb its neighbours."""
    import torch
    imgs = [qoi_test_image(512, 512, 4, 1234), qoi_test_image(300, 500, 3, 6), qoi_test_image(1080, 1920, 4, 7)]
    exp = [oracle.qoi_encode(i) for i in imgs]
    assert codecs.qoi_encode(imgs[0]) == exp[0]
    dev = [torch.from_numpy(i).cuda() for i in imgs]
    outs = [torch.empty(codecs.qoi_encode_bound(i.shape[1], i.shape[0], i.shape[2]) + 16, dtype=torch.uint8, device="cuda") for i in imgs]
    shapes = [i.shape for i in imgs]
    ptrs, optrs = [t.data_ptr() for t in dev], [o.data_ptr() for o in outs]
    lens = codecs.qoi_encode_batch_device(ptrs, shapes, optrs)
    torch.cuda.synchronize()
    for o, n, e in zip(outs, lens, exp):
        assert n == len(e) and o[:n].cpu().numpy().tobytes() == e
    lens = codecs.qoi_encode_batch_device([ptrs[0], ptrs[1], ptrs[2]], [shapes[0], (300, 500, 2), shapes[2]], optrs, pitches=[2048, 1500, 7680])
    assert lens[1] == 0 and lens[0] == len(exp[0]) and lens[2] == len(exp[2])


def test_image_save_qoi(codecs, oracle):
    """Image.saveToMemory(QOI) of a loaded image (saveQOI, plugins/qoi.d:150-185): the stream the reference's encoder
    writes for the image's pixels in whatever layout the image has (gapless, aligned with a border, vertically flipped),
    and it loads back to the same pixels."""
    from gamut_b200.image import Image
    from gamut_b200.types import ImageFormat, PixelType
    from gamut_b200.types import LAYOUT_VERT_FLIPPED, LAYOUT_SCANLINE_ALIGNED_16, LAYOUT_BORDER_2
    for c in (3, 4):
        img = qoi_test_image(37, 61, c, 9)
        src = oracle.qoi_encode(img, colorspace=0)
        for layout in (0, LAYOUT_VERT_FLIPPED, LAYOUT_SCANLINE_ALIGNED_16 | LAYOUT_BORDER_2):
            im = Image()
            assert im.loadFromMemory(src, layout) and im.type() == (PixelType.rgb8 if c == 3 else PixelType.rgba8)
            out = im.saveToMemory(ImageFormat.QOI)
            assert out == src
    im = Image()
    assert im.loadFromMemory(oracle.qoi_encode(qoi_test_image(8, 8, 4, 1)), 0)
    from gamut_b200.types import LOAD_16BIT
    im2 = Image()
    assert im2.loadFromMemory(oracle.qoi_encode(qoi_test_image(8, 8, 4, 1)), LOAD_16BIT)
    assert im2.saveToMemory(ImageFormat.QOI) is None                # saveQOI takes rgb8 / rgba8 only


def test_file_front_end_round_trip(codecs, oracle, tmp_path):
    """Image.loadFromFile / saveToFile (image.d:859-873, 935-958): a QOI file loaded from disk, saved again by extension
    and as TGA by format; the content decides the format when the extension disagrees."""
    from gamut_b200.image import Image
    from gamut_b200.types import ImageFormat
    img = qoi_test_image(20, 30, 4, 5)
    src = oracle.qoi_encode(img)
    p = tmp_path / "a.qoi"
    p.write_bytes(src)
    im = Image()
    assert im.loadFromFile(str(p)) and im.width() == 30 and im.height() == 20
    q = tmp_path / "b.qoi"
    assert im.saveToFile(str(q)) and q.read_bytes() == src
    t = tmp_path / "c.tga"
    assert im.saveToFile(ImageFormat.TGA, str(t)) and t.read_bytes() == oracle.tga_encode(img)
    im2 = Image()
    assert im2.loadFromFile(str(t)) and np.array_equal(im2.pixels(), img)
    w = tmp_path / "d.png"
    w.write_bytes(src)
    assert im2.loadFromFile(str(w)) and im2.width() == 30 and np.array_equal(im2.pixels(), img)
    with open(str(p), "rb") as f:
        im3 = Image()
        assert im3.loadFromStream(f) and np.array_equal(im3.pixels(), img)
