"""Pin the TGA oracle (oracle/tga_oracle.c): PIL's independent reader on what PIL writes, the cited formulas on
hand-made files for the rest (tests/tgautil.py), MemoryFile semantics on truncated files."""
import io

import numpy as np
import pytest

from tgautil import make_tga, pil_tga


def pil_read(data):
    from PIL import Image as PILImage
    return PILImage.open(io.BytesIO(data))


@pytest.mark.parametrize("rle", [False, True])
@pytest.mark.parametrize("top_down", [False, True])
def test_against_pil(oracle, rle, top_down):
    rng = np.random.default_rng(1)
    for c in (1, 3, 4):
        img = rng.integers(0, 4, (23, 37, c)).astype(np.uint8) * 80
        data = pil_tga(img, rle, top_down)
        got = oracle.tga_load(data)
        assert got is not None and np.array_equal(got, img)
        assert np.array_equal(got.reshape(23, 37, c).squeeze(), np.asarray(pil_read(data)))
    from PIL import Image as PILImage
    pimg = PILImage.fromarray(rng.integers(0, 256, (19, 31)).astype(np.uint8), "P")
    pimg.putpalette(bytes(rng.integers(0, 256, 768, dtype=np.uint8)))
    data = pil_tga(pimg, rle, top_down)
    got = oracle.tga_load(data)
    assert got is not None and np.array_equal(got, np.asarray(pil_read(data).convert("RGB")))


def expected_from_source(kind, src, w, h, top_down, data, idlen=0, pal_start=0, pal_len=None, index_bits=8):
    """The decoded image, computed here from the source pixels with the formulas of tga.d (not with the oracle)."""
    def rgb16(v):
        v = v.astype(np.int64)
        return np.stack([((v >> 10) & 31) * 255 // 31, ((v >> 5) & 31) * 255 // 31, (v & 31) * 255 // 31], -1).astype(np.uint8)
    if kind.startswith("pal"):
        bits = int(kind[3:])
        entry = {8: 1, 15: 2, 16: 2, 24: 3, 32: 4}[bits]
        n = pal_len
        pal = np.frombuffer(data, np.uint8, n * entry, 18 + idlen + pal_start).reshape(n, entry)
        idx = src.view("<u2").reshape(-1).astype(np.int64) if index_bits == 16 else src.reshape(-1).astype(np.int64)
        idx = np.where(idx >= n, 0, idx)
        if bits in (15, 16):
            lut = rgb16(pal.view("<u2").reshape(-1))
        elif bits == 8:
            lut = pal
        else:
            lut = pal.copy(); lut[:, [0, 2]] = lut[:, [2, 0]]
        out = lut[idx]
    elif kind in ("rgb15", "rgb16"):
        out = rgb16(src.view("<u2").reshape(-1))
    elif kind in ("bgr24", "bgra32"):
        out = src.copy(); out[:, [0, 2]] = out[:, [2, 0]]
    else:
        out = src
    out = out.reshape(h, w, -1)
    return out if top_down else out[::-1]


CASES = [("l8", {}), ("la16", {}), ("rgb15", {}), ("rgb16", {}), ("bgr24", {"idlen": 7}), ("bgra32", {}),
         ("pal8", {"pal_len": 200}), ("pal15", {"pal_len": 200}), ("pal16", {"pal_len": 200, "pal_start": 5}),
         ("pal24", {"pal_len": 700, "index_bits": 16}), ("pal32", {"pal_len": 90, "idlen": 3})]


@pytest.mark.parametrize("kind,kw", CASES)
@pytest.mark.parametrize("rle", [False, True])
def test_hand_made_variants(oracle, kind, kw, rle):
    rng = np.random.default_rng(7)
    for top_down in (False, True):
        for (w, h) in [(1, 1), (13, 9), (131, 40)]:
            data, src = make_tga(w, h, kind, rng, rle=rle, top_down=top_down, **kw)
            got = oracle.tga_load(data)
            exp = expected_from_source(kind, src, w, h, top_down, data, kw.get("idlen", 0), kw.get("pal_start", 0),
                                       kw.get("pal_len"), kw.get("index_bits", 8))
            assert got is not None and got.shape == exp.shape and np.array_equal(got, exp)


def test_overrun_truncation_and_rejects(oracle):
    rng = np.random.default_rng(3)
    data, src = make_tga(17, 11, "bgr24", rng, rle=True, overrun=True)          # the last packet runs past the image: ignored
    assert np.array_equal(oracle.tga_load(data), expected_from_source("bgr24", src, 17, 11, False, data))
    for kind, rle in [("bgr24", False), ("bgr24", True), ("pal24", False), ("rgb16", True), ("l8", False)]:
        data, _ = make_tga(17, 11, kind, rng, rle=rle, pal_len=50 if kind.startswith("pal") else None)
        assert oracle.tga_load(data) is not None
        assert oracle.tga_load(data[:-1]) is None                                # one byte short: a failed read
        assert oracle.tga_load(data + b"\0\0") is not None                       # trailing bytes are not read
        for cut in (0, 1, 2, 10, 17, 18, 19, 40):
            assert oracle.tga_load(data[:cut]) is None
    ok, _ = make_tga(5, 4, "bgr24", rng)
    bad = bytearray(ok); bad[1] = 2; assert oracle.tga_load(bytes(bad)) is None        # colour-map type > 1
    bad = bytearray(ok); bad[2] = 4; assert oracle.tga_load(bytes(bad)) is None        # image type
    bad = bytearray(ok); bad[16] = 12; assert oracle.tga_load(bytes(bad)) is None      # bits per pixel
    bad = bytearray(ok); bad[12] = bad[13] = 0; assert oracle.tga_load(bytes(bad)) is None   # width 0
    pal, _ = make_tga(5, 4, "pal24", rng, pal_len=9)
    bad = bytearray(pal); bad[5] = bad[6] = 0; assert oracle.tga_load(bytes(bad)) is None    # empty palette
    bad = bytearray(pal); bad[7] = 12; assert oracle.tga_load(bytes(bad)) is None            # palette entry bits
    bad = bytearray(pal); bad[16] = 24; assert oracle.tga_load(bytes(bad)) is None           # index bits


@pytest.mark.parametrize("c", [1, 3, 4])
def test_encoder_read_back_by_pil(oracle, c):
    """or_tga_encode (codecs/tga.d:62-292): PIL's independent reader and the oracle's decoder read the file back to the
    image (l8 is written as a 24-bit file, :91-94); structure checks on the header saveTGA writes (:120-131)."""
    rng = np.random.default_rng(c)
    img = (rng.integers(0, 3, (17, 300, c)) * 100).astype(np.uint8)
    img[5] = 7
    img[6, :200] = 9
    f = oracle.tga_encode(img)
    rgb = img if c >= 3 else np.repeat(img, 3, axis=2)
    assert np.array_equal(np.asarray(pil_read(f)), rgb) and np.array_equal(oracle.tga_load(f), rgb)
    assert f[2] == 10 and f[12] | f[13] << 8 == 300 and f[14] | f[15] << 8 == 17 and f[16] == (24 if c != 4 else 32) and f[17] == 0
    assert len(f) < 18 + img.size * (3 if c == 1 else 1)         # the run-length coding is on
    assert oracle.tga_encode(img, type_=13) is None and oracle.tga_encode(img, shape=(17, 70000, c)) is None
