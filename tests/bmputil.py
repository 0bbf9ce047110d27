"""BMP files for the tests: PIL as an independent writer for the depths it supports, and a small synthesiser for the
header variants it cannot write (OS/2 12-byte header, 4-bit palette, 16-bit 555 / bit fields, 32-bit bit fields with
alpha, V4 / V5 headers, top-down rows, odd data offsets)."""
import io
import struct

import numpy as np
from PIL import Image as PILImage


def pil_bmp(arr, mode=None, **kw):
    im = PILImage.fromarray(arr)
    if mode:
        im = im.convert(mode)
    b = io.BytesIO()
    im.save(b, "BMP", **kw)
    return b.getvalue()


def synth(w, h, bpp, rows, hsz=40, compress=0, masks=None, palette=None, top_down=False, ppm=(0, 0), gap=0, declared_offset=None):
    """rows: h byte strings (already in file order, unpadded). Returns the file."""
    pad = (-len(rows[0])) & 3 if rows else 0
    body = b"".join(r + b"\0" * pad for r in rows)
    pal = b""
    if palette is not None:
        for (r, g, b) in palette:
            pal += bytes((b, g, r)) if hsz == 12 else bytes((b, g, r, 0))
    if hsz == 12:
        hdr = struct.pack("<IHHHH", 12, w, h, 1, bpp)
    else:
        hh = -h if top_down else h
        hdr = struct.pack("<IiiHHIIiiII", hsz, w, hh, 1, bpp, compress, len(body), ppm[0], ppm[1], 0, 0)
        if hsz == 40 and compress == 3:
            hdr += struct.pack("<III", *masks[:3])
        elif hsz == 56:
            hdr += struct.pack("<IIII", *(masks or (0, 0, 0, 0)))
            if compress == 3:
                hdr += struct.pack("<III", *masks[:3])
        elif hsz in (108, 124):
            m = masks or (0, 0, 0, 0)
            hdr += struct.pack("<IIII", *m) + b"\0" * 4 + b"\0" * 48
            if hsz == 124:
                hdr += b"\0" * 16
    off = 14 + len(hdr) + len(pal) + gap
    if declared_offset is not None:
        off = declared_offset
    return b"BM" + struct.pack("<IHHI", off + len(body), 0, 0, off) + hdr + pal + b"\xaa" * gap + body


def variants(rng):
    """(name, file) pairs covering every branch of stbi__bmp_load."""
    out = []
    a = rng.integers(0, 256, (21, 37, 3), dtype=np.uint8)
    al = rng.integers(0, 256, (21, 37, 1), dtype=np.uint8)
    out.append(("pil24", pil_bmp(a)))
    out.append(("pil8", pil_bmp(a, "P")))
    out.append(("pil1", pil_bmp(a, "1")))
    out.append(("pilL", pil_bmp(a, "L")))
    out.append(("pil32", pil_bmp(np.dstack([a, al]))))
    out.append(("pil24dpi", pil_bmp(a, dpi=(200, 100))))
    w, h = 19, 7
    pal16 = [tuple(int(x) for x in rng.integers(0, 256, 3)) for _ in range(16)]
    idx = rng.integers(0, 16, (h, w))
    rows4 = [bytes(((int(r[i]) << 4) | (int(r[i + 1]) if i + 1 < w else 0)) for i in range(0, w, 2)) for r in idx]
    out.append(("4bit", synth(w, h, 4, rows4, palette=pal16)))
    out.append(("4bit_os2", synth(w, h, 4, rows4, hsz=12, palette=pal16)))
    out.append(("4bit_topdown", synth(w, h, 4, rows4, palette=pal16, top_down=True)))
    out.append(("4bit_gap", synth(w, h, 4, rows4, palette=pal16, gap=5)))
    pal2 = [(0, 0, 0), (255, 200, 10)]
    bits = rng.integers(0, 2, (h, w))
    rows1 = [bytes(sum(int(r[i + k]) << (7 - k) for k in range(8) if i + k < w) for i in range(0, w, 8)) for r in bits]
    out.append(("1bit_os2", synth(w, h, 1, rows1, hsz=12, palette=pal2)))
    v16 = rng.integers(0, 65536, (h, w), dtype=np.uint16)
    rows16 = [r.astype("<u2").tobytes() for r in v16]
    out.append(("16bit_555", synth(w, h, 16, rows16)))
    out.append(("16bit_565", synth(w, h, 16, rows16, compress=3, masks=(0xF800, 0x07E0, 0x001F))))
    out.append(("16bit_v4_4444", synth(w, h, 16, rows16, hsz=108, compress=3, masks=(0x0F00, 0x00F0, 0x000F, 0xF000))))
    v32 = rng.integers(0, 2 ** 32, (h, w), dtype=np.uint64).astype(np.uint32)
    rows32 = [r.astype("<u4").tobytes() for r in v32]
    out.append(("32bit_rgb", synth(w, h, 32, rows32)))
    out.append(("32bit_alpha0", synth(w, h, 32, [(r & 0x00FFFFFF).astype("<u4").tobytes() for r in v32])))
    out.append(("32bit_fields", synth(w, h, 32, rows32, compress=3, masks=(0x0000FF00, 0x00FF0000, 0xFF000000))))
    out.append(("32bit_fields_565", synth(w, h, 32, rows32, compress=3, masks=(0x001F0000, 0x000007E0, 0xF8000000))))
    out.append(("32bit_v5_fields", synth(w, h, 32, rows32, hsz=124, compress=3, masks=(0x00FF0000, 0x0000FF00, 0x000000FF, 0xFF000000))))
    out.append(("32bit_v4_rgb", synth(w, h, 32, rows32, hsz=108)))
    out.append(("32bit_56", synth(w, h, 32, rows32, hsz=56)))
    rows24 = [rng.integers(0, 256, w * 3, dtype=np.uint8).tobytes() for _ in range(h)]
    out.append(("24bit_topdown", synth(w, h, 24, rows24, top_down=True, ppm=(7874, 3937))))
    out.append(("24bit_os2", synth(w, h, 24, rows24, hsz=12)))
    return out


def broken(rng):
    """Files that must fail or decode with zero padding: same decision as the oracle is what the tests check."""
    good = dict(variants(rng))
    f24, f8, f32 = good["pil24"], good["pil8"], good["32bit_rgb"]
    out = [("truncated_rows", f24[:len(f24) // 2]), ("truncated_header", f24[:30]), ("only_magic", b"BM"), ("rle8", f8[:30] + struct.pack("<I", 1) + f8[34:]),
           ("hsz52", f24[:14] + struct.pack("<I", 52) + f24[18:]), ("planes2", f24[:26] + struct.pack("<H", 2) + f24[28:]),
           ("bpp2", f8[:28] + struct.pack("<H", 2) + f8[30:]), ("offset_small", f24[:10] + struct.pack("<I", 20) + f24[14:]),
           ("offset_huge", f24[:10] + struct.pack("<I", 5000) + f24[14:]), ("offset_negative", f24[:10] + struct.pack("<i", -5) + f24[14:]),
           ("fields_equal", synth(5, 3, 32, [b"\1\2\3\4" * 5] * 3, compress=3, masks=(0xFF, 0xFF, 0xFF))),
           ("fields_wide", synth(5, 3, 32, [b"\1\2\3\4" * 5] * 3, compress=3, masks=(0x3FF00000, 0x000FFC00, 0x000003FF))),
           ("pal_offset_before_palette", f8[:10] + struct.pack("<I", 60) + f8[14:]),
           ("truncated_palette", f8[:70]), ("trunc32", f32[:len(f32) - 17])]
    return out
