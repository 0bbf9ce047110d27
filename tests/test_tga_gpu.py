"""GPU parity of the TGA decoder: gb200_tga_load / gb200_tga_decode_batch / Image.loadFromMemory against
the oracle's restatement of TGADecoder (codecs/tga.d:313-588, pinned to PIL and to the format's formulas in
tests/test_oracle_tga.py) on every variant of tests/tgautil.py: grey, grey + alpha, 15/16/24/32-bit colour, palettes
with 8- and 16-bit indices, raw and run-length coded, both orientations, ID field, truncated files."""
import numpy as np
import pytest

from tgautil import make_tga, pil_tga
from test_oracle_tga import CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codecs(gb):
    from gamut_b200 import codecs
    return codecs


def same(a, b):
    return (a is None and b is None) or (a is not None and b is not None and a.shape == b.shape and np.array_equal(a, b))


@pytest.mark.parametrize("kind,kw", CASES)
def test_variants(codecs, oracle, kind, kw):
    rng = np.random.default_rng(7)
    for rle in (False, True):
        for top_down in (False, True):
            for (w, h) in [(1, 1), (13, 9), (131, 40), (640, 33)]:
                data, _ = make_tga(w, h, kind, rng, rle=rle, top_down=top_down, **kw)
                exp = oracle.tga_load(data)
                assert exp is not None and same(codecs.tga_load(data), exp)
            for cut in (len(data) - 1, len(data) // 2, 19, 18, 5, 0):
                assert same(codecs.tga_load(data[:cut]), oracle.tga_load(data[:cut]))


def test_pil_files_large_and_rejects(codecs, oracle):
    rng = np.random.default_rng(1)
    for c in (1, 3, 4):
        img = rng.integers(0, 4, (23, 37, c)).astype(np.uint8) * 80
        for rle in (False, True):
            assert same(codecs.tga_load(pil_tga(img, rle, bool(c & 1))), img)
    big = np.zeros((1080, 1920, 4), np.uint8)                    # runs and noise at a real size, both codings
    big[..., :3] = np.linspace(0, 255, 1920)[None, :, None].astype(np.uint8)
    big[200:500, 300:900] = rng.integers(0, 256, (300, 600, 4))
    big[..., 3] = 255
    for rle in (False, True):
        assert same(codecs.tga_load(pil_tga(big, rle)), big)
    data, _ = make_tga(17, 11, "bgr24", rng, rle=True, overrun=True)
    assert same(codecs.tga_load(data), oracle.tga_load(data))
    ok, _ = make_tga(5, 4, "pal24", rng, pal_len=9)
    for _ in range(100):                                         # header fuzz: both sides accept / reject the same files
        bad = bytearray(ok)
        for _ in range(int(rng.integers(1, 4))):
            bad[int(rng.integers(0, 18))] = int(rng.integers(0, 256))
        bad[12:16] = (5).to_bytes(2, "little") + (4).to_bytes(2, "little")
        assert same(codecs.tga_load(bytes(bad)), oracle.tga_load(bytes(bad)))


def test_batch(codecs, oracle):
    rng = np.random.default_rng(5)
    files = []
    for kind, kw in CASES:
        for rle in (False, True):
            files.append(make_tga(57, 23, kind, rng, rle=rle, top_down=bool(len(files) & 1), **kw)[0])
    files += [b"nope", files[3][:-1], files[0][:40]]
    b = codecs.tga_decode_batch(files)
    try:
        for i, f in enumerate(files):
            exp = oracle.tga_load(f)
            got = b.to_host(i)
            if exp is None:
                assert got is None and b.images[i].status == 0
            else:
                assert np.array_equal(got.reshape(exp.shape), exp)
    finally:
        b.free()


def test_image_load(codecs, oracle):
    """Image.loadFromMemory on TGA files (plugins/tga.d:45-105 -> convertTo): type, pitch and every scanline equal to the
    oracle-side restatement of the plugin for a few flag / layout combinations; the staged path agrees."""
    from gamut_b200.image import Image, kStrImageDecodingFailed
    from gamut_b200.types import (ImageFormat, PixelType as PT, LOAD_RGB, LOAD_ALPHA, LOAD_GREYSCALE, LOAD_FP32, LOAD_16BIT,
                                  LAYOUT_VERT_FLIPPED, LAYOUT_SCANLINE_ALIGNED_16, LAYOUT_BORDER_1)
    from oracle import pyimage
    rng = np.random.default_rng(2)
    for kind, kw, rle in [("bgr24", {}, True), ("bgra32", {}, False), ("l8", {}, True), ("la16", {}, False), ("pal24", {"pal_len": 99}, True), ("rgb16", {}, False)]:
        data, _ = make_tga(45, 31, kind, rng, rle=rle, **kw)
        assert Image.identifyFormatFromMemory(data) == ImageFormat.TGA
        for flags in (0, LOAD_RGB | LOAD_ALPHA, LOAD_GREYSCALE | LOAD_FP32, LOAD_16BIT | LAYOUT_VERT_FLIPPED, LAYOUT_SCANLINE_ALIGNED_16 | LAYOUT_BORDER_1):
            exp = pyimage.load_from_memory(data, flags)
            im = Image()
            im.loadFromMemory(data, flags)
            assert (exp.error is None) == im.isValid()
            if exp.error is None:
                assert int(im.type()) == exp.type and im.pitchInBytes() == exp.pitch and im.width() == 45 and im.height() == 31
                for y in range(im.height()):
                    assert np.array_equal(im.scanline(y), exp.scanline(y))
            im2 = Image()
            im2.loadFromMemoryStaged(data, flags)
            assert im2.isValid() == im.isValid()
            if im.isValid():
                assert int(im2.type()) == int(im.type())
                for y in range(im.height()):
                    assert np.array_equal(im2.scanline(y), im.scanline(y))
    im = Image()
    data, _ = make_tga(45, 31, "bgr24", rng)
    im.loadFromMemory(data[:-5])
    assert im.isError() and im.errorMessage() == kStrImageDecodingFailed


# ---- encoder: saveTGA (plugins/tga.d:123-149) -> TGAEncoder (codecs/tga.d:62-292) ---------------------------------------
@pytest.mark.parametrize("c", [1, 2, 3, 4])
def test_encoder_files_equal_the_oracle(codecs, oracle, c):
    from test_tga_emulated import tga_encode_images
    rng = np.random.default_rng(20 + c)
    imgs = tga_encode_images(c, rng)
    big = np.zeros((1080, 1920, c), np.uint8)
    big[..., 0] = np.linspace(0, 255, 1920)[None, :].astype(np.uint8)
    big[200:500, 300:900] = rng.integers(0, 256, (300, 600, c))
    imgs.append(big)
    for img in imgs:
        exp = oracle.tga_encode(img)
        got = codecs.tga_encode(img)
        assert exp is not None and got is not None and len(got) == len(exp) and got == exp
        rgb = img if c >= 3 else np.concatenate([np.repeat(img[..., :1], 3, axis=2), img[..., 1:]], axis=2)
        assert np.array_equal(codecs.tga_load(got), rgb)          # both decoders read it back
        assert np.array_equal(oracle.tga_load(got), rgb)


def test_encoder_pitch_flip_batch_and_rejects(codecs, oracle):
    import torch
    rng = np.random.default_rng(4)
    img = (rng.integers(0, 3, (21, 45, 4)) * 90).astype(np.uint8)
    exp = oracle.tga_encode(img)
    wide = rng.integers(0, 256, (21, 60, 4)).astype(np.uint8)
    wide[:, :45] = img
    assert codecs.tga_encode(wide, pitch=240, shape=(21, 45, 4)) == exp
    flipped = np.ascontiguousarray(wide[::-1])
    assert codecs.tga_encode(flipped, pitch=-240, first_scanline=20 * 240, shape=(21, 45, 4)) == exp
    # saveTGA's refusals: pixel types other than l8 / la8 / rgb8 / rgba8; and a pitch smaller than a scanline
    assert codecs.tga_encode(img, type_=13) is None and codecs.tga_encode(img, type_=1) is None
    assert codecs.tga_encode(img, pitch=100) is None
    assert codecs.tga_encode(np.zeros((0, 5, 3), np.uint8) if False else img[:0], shape=(0, 45, 4)) == oracle.tga_encode(img, shape=(0, 45, 4))
    # one device batch of mixed types; a refused image in the middle does not disturb its neighbours
    imgs = [(rng.integers(0, 3, (300, 500, 3)) * 100).astype(np.uint8), rng.integers(0, 256, (64, 64, 1)).astype(np.uint8),
            (rng.integers(0, 2, (512, 512, 4)) * 255).astype(np.uint8)]
    exps = [oracle.tga_encode(i) for i in imgs]
    dev = [torch.from_numpy(i).cuda() for i in imgs]
    outs = [torch.empty(codecs.tga_encode_bound(i.shape[1], i.shape[0], i.shape[2]), dtype=torch.uint8, device="cuda") for i in imgs]
    lens = codecs.tga_encode_batch_device([t.data_ptr() for t in dev], [i.shape for i in imgs], [o.data_ptr() for o in outs])
    torch.cuda.synchronize()
    for o, k, e in zip(outs, lens, exps):
        assert k == len(e) and o[:k].cpu().numpy().tobytes() == e
    lens = codecs.tga_encode_batch_device([t.data_ptr() for t in dev], [imgs[0].shape, (64, 64, 5), imgs[2].shape], [o.data_ptr() for o in outs])
    assert lens[1] == 0 and lens[0] == len(exps[0]) and lens[2] == len(exps[2])


def test_image_save_tga(codecs, oracle):
    """Image.saveToMemory(TGA): load a TGA into several layouts, save it again: the file saveTGA writes for those pixels."""
    from gamut_b200.image import Image
    from gamut_b200.types import ImageFormat, LAYOUT_VERT_FLIPPED, LAYOUT_SCANLINE_ALIGNED_16, LAYOUT_BORDER_2, LOAD_16BIT
    rng = np.random.default_rng(8)
    for c in (1, 2, 3, 4):
        img = (rng.integers(0, 3, (37, 61, c)) * 100).astype(np.uint8)
        kind = {1: "l8", 2: "la16", 3: "bgr24", 4: "bgra32"}[c]
        exp = oracle.tga_encode(img)
        src = pil_tga(img, True) if c != 2 else None
        if src is None:                                          # PIL cannot write grey + alpha: a raw hand-made file
            from tgautil import header
            src = header(0, 0, 3, 0, 0, 0, 61, 37, 16, 0x20) + img.tobytes()
        for layout in (0, LAYOUT_VERT_FLIPPED, LAYOUT_SCANLINE_ALIGNED_16 | LAYOUT_BORDER_2):
            im = Image()
            assert im.loadFromMemory(src, layout) and im.width() == 61
            assert im.saveToMemory(ImageFormat.TGA) == exp
    im = Image()
    assert im.loadFromMemory(pil_tga((rng.integers(0, 3, (8, 8, 3)) * 100).astype(np.uint8)), LOAD_16BIT)
    assert im.saveToMemory(ImageFormat.TGA) is None              # rgb16: unsupported by TGAEncoder.initialize
