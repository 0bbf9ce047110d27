"""BMP oracle (oracle/bmp_oracle.c, a restatement of stbi__bmp_load, stbdec.d:2112-2510) and the format detection
(image.d:1045-1061) against the reference's own KAT and independent decoders."""
import io
import os

import numpy as np
from PIL import Image as PILImage

from bmputil import variants, broken

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_issue67_kat(oracle):
    """examples/test-suite/source/main.d:161-170: 200 x 100 dpi, pixel aspect ratio 2; :194-197: 32 x 32."""
    data = open(os.path.join(G, "issue67.bmp"), "rb").read()
    px, comp, ppmX, ppmY, ratio = oracle.bmp_load(data, 0)
    assert px.shape[:2] == (32, 32)
    resY = np.float32(ppmY) / np.float32(39.37007874)              # plugins/bmp.d:134
    assert abs(resY - 100) < 0.1 and abs(resY * ratio - 200) < 0.1 and abs(ratio - 2) < 0.01
    from oracle import pyimage
    im = pyimage.load_from_memory(data, 0)
    assert im.error is None and abs(im.resY - 100) < 0.1 and abs(im.par - 2) < 0.01


def test_against_pil(oracle):
    """Every variant PIL can read, pixel-exact (PIL drops the alpha byte of 32-bit BI_RGB files: compare RGB)."""
    rng = np.random.default_rng(5)
    n = 0
    for name, f in variants(rng):
        r = oracle.bmp_load(f, 0)
        assert r is not None, name
        try:
            ref = np.asarray(PILImage.open(io.BytesIO(f)).convert("RGB"))
        except Exception:
            continue
        if name.startswith("16bit") or name in ("32bit_fields_565",):
            continue                    # PIL expands 5/6/10-bit fields differently (stb: stbi__shiftsigned replication)
        if name in ("4bit_os2", "1bit_os2"):
            continue                    # stb sizes an OS/2 palette as (offset - 14 - 24) / 3 (stbdec.d:2288): 4 entries short
        assert np.array_equal(r[0][:, :, :3], ref), name
        n += 1
    assert n >= 11


def test_field_expansion(oracle):
    """stbi__shiftsigned (stbdec.d:2493-2512): an n-bit field is replicated to 8 bits (v * mul >> shift)."""
    from bmputil import synth
    for bits, mask, shift in ((5, 0x7C00, 10), (6, 0x07E0, 5), (4, 0x0F00, 8)):
        for v in range(1 << bits):
            word = v << shift
            masks = (mask, 0x001F if mask != 0x001F else 0x7C00, 0x0001 if bits != 1 else 2)
            f = synth(1, 1, 16, [int(word).to_bytes(2, "little")], compress=3, masks=(mask, 0x8000 if not (mask & 0x8000) else 0x0001, 0x0010 if not (mask & 0x0010) else 0x0002))
            r = oracle.bmp_load(f, 0)
            assert r is not None
            exp = v
            b = bits
            while b < 8:
                exp = (exp << bits) | v
                b += bits
            exp >>= (b - 8)
            assert int(r[0][0, 0, 0]) == exp, (bits, v)


def test_req_comp_and_alpha_rule(oracle):
    rng = np.random.default_rng(6)
    v = dict(variants(rng))
    a32 = oracle.bmp_load(v["32bit_rgb"], 0)[0]
    assert a32.shape[2] == 4
    z = oracle.bmp_load(v["32bit_alpha0"], 0)[0]
    assert (z[:, :, 3] == 255).all()                     # all-zero alpha is replaced by 255 (stbdec.d:2438-2443)
    for req in (1, 2, 3, 4):
        r = oracle.bmp_load(v["32bit_rgb"], req)[0]
        assert r.shape[2] == req
        if req >= 3:
            assert np.array_equal(r[:, :, :3], a32[:, :, :3])
        else:
            y = ((a32[:, :, 0].astype(int) * 77 + a32[:, :, 1].astype(int) * 150 + a32[:, :, 2].astype(int) * 29) >> 8).astype(np.uint8)
            assert np.array_equal(r[:, :, 0], y)
            if req == 2:
                assert np.array_equal(r[:, :, 1], a32[:, :, 3])


def test_broken_files_do_not_crash(oracle):
    rng = np.random.default_rng(7)
    res = {name: oracle.bmp_load(f, 0) for name, f in broken(rng)}
    for name in ("only_magic", "rle8", "hsz52", "planes2", "bpp2", "offset_small", "offset_huge", "offset_negative", "fields_equal", "fields_wide"):
        assert res[name] is None, name
    assert res["truncated_rows"] is not None and (res["truncated_rows"][0][0] == 0).all()      # missing rows read as 0 (bottom-up: they are the top)


def test_identify_format(oracle):
    """Image.identifyFormatFromStream (image.d:1045-1061): ImageFormat order, TGA last."""
    F = oracle.identify_format
    assert F(open(os.path.join(G, "issue35.jpg"), "rb").read()) == 0
    assert F(open(os.path.join(G, "issue76.png"), "rb").read()) == 1
    assert F(b"qoif" + b"\0" * 20) == 2 and F(b"qoix" + b"\0" * 30) == 3 and F(b"DDS " + b"\0" * 20) == 4
    assert F(b"GIF89a" + b"\0" * 20) == 6 and F(b"GIF87a") == 6 and F(b"GIF88a" + b"\0" * 30) == -1
    assert F(open(os.path.join(G, "issue67.bmp"), "rb").read()) == 7
    assert F(b"BM" + b"\0" * 12 + (52).to_bytes(4, "little")) == 7 and F(b"BM" + b"\0" * 12 + (53).to_bytes(4, "little")) == -1
    assert F(b"\xff\x0a") == 8 and F(b"\xa5") == 9 and F(b"") == -1 and F(b"nope") == -1
    tga = bytes([0, 0, 2]) + b"\0" * 9 + (4).to_bytes(2, "little") + (4).to_bytes(2, "little") + bytes([24, 0])
    assert F(tga) == 5 and F(tga[:16]) == -1
    assert F(bytes([0, 0, 2]) + b"\0" * 9 + (0).to_bytes(2, "little") + (4).to_bytes(2, "little") + bytes([24, 0])) == -1
