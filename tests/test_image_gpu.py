"""GPU parity of the plugin epilogues and Image.convertTo: gamut_b200.Image.loadFromMemory(bytes, flags) (decode and
conversion on the GPU through the C ABI) against the CPU oracle's independent restatement (oracle/pyimage.py), on the
reference's own scenarios:

    examples/test-suite/source/main.d:28-35    issue35.jpg, LOAD_RGB|LOAD_8BIT|LOAD_ALPHA|VERT_STRAIGHT|GAPLESS
    examples/test-suite/source/main.d:38-49    issue46.jpg fails, the Image stays usable
    examples/test-suite/source/main.d:137-159  issue65.png, LOAD_FP32|LOAD_GREYSCALE, setLayout x2, convertTo8Bit
    examples/test-suite/source/main.d:165-184  issue76.png -> l16 [[1875, 65535], [0, 2807]]
    examples/test-suite/source/main.d:215-221  vst3-compatible.png -> rgb8, VERT_FLIPPED | BORDER_3
    source/gamut/image.d:2112-2183             3x1 rgb8 round trip through PNG / QOI / QOIX, loaded then convertTo(rgb8)

plus flipped / bordered / aligned / trailing layouts and every LoadFlags family on all four formats. Compared: error
string, PixelType, width, height, |pitch| and its sign, every scanline's bytes, resolution fields."""
import io
import os

import numpy as np
import pytest

from gamut_b200.types import *  # noqa: F401,F403
from gamut_b200.types import PixelType as PT

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return open(os.path.join(G, name), "rb").read()


@pytest.fixture(scope="module")
def Image(gb):
    from gamut_b200.image import Image
    return Image


@pytest.fixture(scope="module")
def pyimage(oracle):
    from oracle import pyimage
    return pyimage


def same(im, o, what=""):
    """gamut_b200.Image vs oracle OImage."""
    if o.error:
        assert im.isError() and im.errorMessage() == o.error, (what, im.errorMessage(), o.error)
        return
    assert im.isValid(), (what, im.errorMessage())
    assert int(im.type()) == o.type and im.width() == o.w and im.height() == o.h, (what, im.type(), o.type)
    assert im.pitchInBytes() == o.pitch, (what, im.pitchInBytes(), o.pitch)
    assert im.layoutConstraints() == o.layout, what
    for y in range(o.h):
        assert np.array_equal(im.scanline(y), o.scanline(y)), (what, y)
    a, b = np.float32(im._pixelAspectRatio), np.float32(o.par)
    assert a == b or (np.isnan(a) and np.isnan(b)), what
    a, b = np.float32(im._resolutionY), np.float32(o.resY)
    assert a == b or (np.isnan(a) and np.isnan(b)), what


def load_both(Image, pyimage, data, flags):
    im = Image()
    im.loadFromMemory(data, flags)
    return im, pyimage.load_from_memory(data, flags)


def test_issue35_jpeg(Image, pyimage):
    f = LOAD_RGB | LOAD_8BIT | LOAD_ALPHA | LAYOUT_VERT_STRAIGHT | LAYOUT_GAPLESS
    im, o = load_both(Image, pyimage, gold("issue35.jpg"), f)
    assert not im.isError() and im._layerCount == 1 and im.type() == PT.rgba8
    same(im, o, "issue35")


def test_issue46_empty_jpeg_then_reuse(Image, pyimage):
    im = Image()
    im.loadFromMemory(gold("issue46.jpg"))
    assert im.isError()
    im.loadFromMemory(gold("issue35.jpg"))
    assert not im.isError()
    same(im, pyimage.load_from_memory(gold("issue35.jpg"), 0), "issue35 default flags")
    im.loadFromMemory(gold("issue46.jpg"))
    assert im.isError()
    assert pyimage.load_from_memory(gold("issue46.jpg"), 0).error == im.errorMessage()


def test_issue65_fp32_grey_then_layouts(Image, pyimage):
    data = gold("issue65.png")
    im, o = load_both(Image, pyimage, data, LOAD_FP32 | LOAD_GREYSCALE)
    assert im.hasData() and im.isValid()
    same(im, o, "issue65 load")
    for step, lay in (("trailing1", LAYOUT_TRAILING_1), ("trailing0", LAYOUT_TRAILING_0)):
        assert im.setLayout(lay) and o.convert_to(o.type, lay), step
        assert im.hasData()
        same(im, o, step)
    assert im.convertTo8Bit() and o.convert_to(o.type - o.type % 3, 0)
    assert im.hasData()
    same(im, o, "to 8 bit")


def test_issue76_l16_kat(Image, pyimage):
    im, o = load_both(Image, pyimage, gold("issue76.png"), 0)
    assert im.isValid() and im.type() == PT.l16 and im.width() == 2 and im.height() == 2
    s0, s1 = im.scanline(0).view(np.uint16), im.scanline(1).view(np.uint16)
    assert s0.tolist() == [1875, 65535] and s1.tolist() == [0, 2807]          # main.d:178-183
    same(im, o, "issue76")


def test_issue77_flipped_border3(Image, pyimage):
    data = gold("vst3-compatible.png")
    im, o = load_both(Image, pyimage, data, 0)
    same(im, o, "vst3 load")
    lay = LAYOUT_VERT_FLIPPED | LAYOUT_BORDER_3
    assert im.convertTo(PT.rgb8, lay) and o.convert_to(int(PT.rgb8), lay)
    assert im.pitchInBytes() < 0
    same(im, o, "vst3 rgb8 flipped border 3")


@pytest.mark.parametrize("fmt", ["png", "qoi", "qoix"])
def test_3x1_roundtrip_kat(Image, pyimage, oracle, fmt):
    """image.d:2112-2183: [255,0,0, 15,64,255, 0,255,255] survives every lossless codec, loaded with default flags and
    converted to rgb8. Encoders: PIL (PNG, QOI -- independent implementations) and tests/qoixsynth.py (QOIX: 8-bit rgb
    is the QOI2AVG sub-codec; literal opcodes written from the format description)."""
    from PIL import Image as PILImage
    px = np.array([[[255, 0, 0], [15, 64, 255], [0, 255, 255]]], np.uint8)
    if fmt == "qoix":
        import qoixsynth                      # 8-bit RGB QOIX = QOI2AVG; its stream is written by the independent synthesiser
        data = qoixsynth.encode_qoi2avg(px)
    else:
        b = io.BytesIO()
        PILImage.fromarray(px, "RGB").save(b, fmt.upper())
        data = b.getvalue()
    im, o = load_both(Image, pyimage, data, 0)
    assert im.convertTo(PT.rgb8) and o.convert_to(int(PT.rgb8), 0)
    assert not im.isError() and im._layerCount == 1 and im.width() == 3 and im.height() == 1
    assert im.scanline(0).tolist() == [255, 0, 0, 15, 64, 255, 0, 255, 255]
    same(im, o, fmt)


def _files():
    """One small file per format (and per interesting variant)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from PIL import Image as PILImage
    from oracle import pyoracle
    from qoixutil import depth_map_la, qoi_bytes, qoi_test_image
    import jpegutil
    from pngwriter import write_png
    rng = np.random.default_rng(5)
    out = {}
    out["png_rgba8"] = write_png(rng.integers(0, 256, (13, 21, 4)).astype(np.uint8), 6, 8, filters=(0, 1, 2, 3, 4))
    out["png_la16"] = write_png(rng.integers(0, 65536, (9, 14, 2)), 4, 16, filters=4, phys=(3780, 3780, 1))
    out["png_l8"] = write_png(rng.integers(0, 256, (7, 8, 1)).astype(np.uint8), 0, 8, filters=3)
    out["jpeg_420"] = jpegutil.encode(jpegutil.photo(40, 56, 3, 1), 90, 2, dpi=(72, 96))
    out["jpeg_grey"] = jpegutil.encode(jpegutil.photo(24, 33, 1, 2), 85, 0)
    out["qoi_rgba"] = qoi_bytes(qoi_test_image(16, 24, 4, 3))
    out["qoi_rgb"] = qoi_bytes(qoi_test_image(9, 17, 3, 4))
    out["qoix_la10"] = pyoracle.qoix_encode(depth_map_la(18, 25, 6, 2), 10, force_lz4=True, par=1.5, dpi=96.0)
    out["qoix_l10"] = pyoracle.qoix_encode(depth_map_la(11, 16, 7, 1), 10)
    return out


FLAG_SETS = [0, LOAD_GREYSCALE, LOAD_RGB, LOAD_ALPHA, LOAD_NO_ALPHA, LOAD_GREYSCALE | LOAD_ALPHA, LOAD_GREYSCALE | LOAD_NO_ALPHA,
             LOAD_RGB | LOAD_ALPHA, LOAD_RGB | LOAD_NO_ALPHA, LOAD_8BIT, LOAD_16BIT, LOAD_FP32, LOAD_PREMUL | LOAD_ALPHA,
             LOAD_NO_PREMUL, LOAD_RGB | LOAD_ALPHA | LOAD_FP32 | LOAD_PREMUL, LOAD_GREYSCALE | LOAD_16BIT,
             LOAD_GREYSCALE | LOAD_RGB, LOAD_ALPHA | LOAD_NO_ALPHA, LOAD_8BIT | LOAD_FP32]
LAYOUTS = [0, LAYOUT_GAPLESS | LAYOUT_VERT_STRAIGHT, LAYOUT_VERT_FLIPPED, LAYOUT_VERT_FLIPPED | LAYOUT_GAPLESS,
           LAYOUT_BORDER_2 | LAYOUT_TRAILING_3, LAYOUT_SCANLINE_ALIGNED_64, LAYOUT_MULTIPLICITY_8 | LAYOUT_TRAILING_7 | LAYOUT_SCANLINE_ALIGNED_16,
           LAYOUT_VERT_FLIPPED | LAYOUT_BORDER_1 | LAYOUT_MULTIPLICITY_4]


def test_every_flag_family_on_every_format(Image, pyimage):
    files = _files()
    n = 0
    for name, data in files.items():
        for f in FLAG_SETS:
            im, o = load_both(Image, pyimage, data, f)
            same(im, o, (name, hex(f)))
            n += 1
    assert n == len(files) * len(FLAG_SETS)


def test_layouts_on_every_format(Image, pyimage):
    files = _files()
    for name, data in files.items():
        for lay in LAYOUTS:
            for f in (0, LOAD_RGB | LOAD_ALPHA | LOAD_8BIT, LOAD_FP32):
                im, o = load_both(Image, pyimage, data, f | lay)
                same(im, o, (name, hex(f), lay))
                if o.error is None and (lay & LAYOUT_VERT_FLIPPED) and o.h >= 2:
                    assert im.pitchInBytes() < 0
                if o.error is None:
                    al = 1 << ((lay >> 4) & 0x0F)
                    assert (im._area.ctypes.data + im._offset) % al == 0 and im.pitchInBytes() % al == 0


def test_convert_after_load_chain(Image, pyimage):
    """convertTo on an already constrained image: the source has a negative pitch, a border and padding."""
    files = _files()
    for name in ("png_rgba8", "jpeg_420", "qoix_la10"):
        im, o = load_both(Image, pyimage, files[name], LAYOUT_VERT_FLIPPED | LAYOUT_BORDER_2 | LAYOUT_TRAILING_1)
        same(im, o, name)
        for t, lay in ((PT.rgbaf32, LAYOUT_SCANLINE_ALIGNED_32), (PT.la16, LAYOUT_VERT_FLIPPED), (PT.rgb8, LAYOUT_GAPLESS),
                       (PT.rgbap8, LAYOUT_BORDER_3), (PT.l8, 0)):
            assert im.convertTo(t, lay) == o.convert_to(int(t), lay)
            same(im, o, (name, t.name, lay))


def test_single_pass_load_equals_staged_load(Image):
    """gb200_image_load (decode + convert on the GPU into the final layout, one copy back) against the same load
    composed from the codec entry points + convertTo, the way the reference's plugins do it."""
    files = _files()
    for name, data in files.items():
        for f in FLAG_SETS[:12]:
            for lay in (0, LAYOUT_VERT_FLIPPED | LAYOUT_BORDER_1, LAYOUT_SCANLINE_ALIGNED_32 | LAYOUT_TRAILING_3):
                a, b = Image(), Image()
                a.loadFromMemory(data, f | lay)
                b.loadFromMemoryStaged(data, f | lay)
                assert a.isError() == b.isError() and a.errorMessage() == b.errorMessage(), (name, hex(f), lay)
                if a.isError():
                    continue
                assert (a.type(), a.width(), a.height(), a.pitchInBytes(), a.layoutConstraints()) == \
                       (b.type(), b.width(), b.height(), b.pitchInBytes(), b.layoutConstraints()), (name, hex(f), lay)
                for y in range(a.height()):
                    assert np.array_equal(a.scanline(y), b.scanline(y)), (name, hex(f), lay, y)


def test_unidentified_and_garbage(Image, pyimage):
    for data in (b"", b"abc", b"\x89PNG\r\n\x1a\n", b"qoif", b"qoix" + b"\0" * 30, b"\xff\xd8\xff"):
        im, o = load_both(Image, pyimage, data, 0)
        assert im.isError()
        assert im.errorMessage() == o.error, data
