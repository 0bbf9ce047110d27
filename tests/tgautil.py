"""Hand-made TGA files for the variants PIL cannot write (15/16-bit, 16-bit indices, grey + alpha, ID field, palette
start, packets that cross rows, truncation), and PIL-written ones for the rest. Format: the TARGA header that
TGADecoder.getImageInfo / decodeImage read (codecs/tga.d:313-420)."""
import io
import struct

import numpy as np


def header(idlen, cmap_type, image_type, pal_start, pal_len, cmap_bits, w, h, bpp, descriptor):
    return struct.pack("<BBBHHBHHHHBB", idlen, cmap_type, image_type, pal_start, pal_len, cmap_bits, 0, 0, w, h, bpp, descriptor)


def rle_packets(pixels, rng, bpp_bytes):
    """pixels: (n, bpp_bytes) uint8 source pixels in file order -> RLE packet stream (random mix of run / raw packets)."""
    out = bytearray()
    i, n = 0, len(pixels)
    while i < n:
        run = 1
        while i + run < n and run < 128 and (pixels[i + run] == pixels[i]).all():
            run += 1
        if run > 1 and rng.integers(0, 4) > 0:
            k = int(rng.integers(1, run + 1))
            out.append(0x80 | (k - 1)); out += pixels[i].tobytes()
            i += k
        else:
            k = int(min(n - i, rng.integers(1, 129)))
            out.append(k - 1); out += pixels[i:i + k].tobytes()
            i += k
    return bytes(out)


def make_tga(w, h, kind, rng, rle=False, top_down=False, idlen=0, pal_start=0, pal_len=None, index_bits=8, overrun=False):
    """kind: 'l8', 'la16' (grey + alpha), 'rgb15', 'rgb16', 'bgr24', 'bgra32', 'pal8' / 'pal15' / 'pal16' / 'pal24' / 'pal32'
    (palette entry bits). Returns (file bytes, source pixels (h*w, bytes per source pixel) in file order)."""
    desc = 0x20 if top_down else 0
    ident = bytes(rng.integers(0, 256, idlen, dtype=np.uint8))
    pal = b""
    if kind.startswith("pal"):
        bits = int(kind[3:])
        n = pal_len if pal_len is not None else (200 if index_bits == 8 else 700)
        entry = {8: 1, 15: 2, 16: 2, 24: 3, 32: 4}[bits]
        pal = bytes(rng.integers(0, 256, pal_start, dtype=np.uint8)) + bytes(rng.integers(0, 256, n * entry, dtype=np.uint8))
        hi = min(n + 20, 256 if index_bits == 8 else 65536)                     # some indices past the palette -> entry 0
        idx = rng.integers(0, hi, w * h)
        idx[rng.integers(0, w * h, w * h // 3)] = idx[0]                         # runs
        idx = np.sort(idx.reshape(h, w), axis=1).reshape(-1) if rle else idx
        src = idx.astype("<u2").view(np.uint8).reshape(-1, 2) if index_bits == 16 else idx.astype(np.uint8).reshape(-1, 1)
        hdr = header(idlen, 1, 9 if rle else 1, pal_start, n, bits, w, h, index_bits, desc)
    else:
        nb, itype = {"l8": (1, 3), "la16": (2, 3), "rgb15": (2, 2), "rgb16": (2, 2), "bgr24": (3, 2), "bgra32": (4, 2)}[kind]
        bpp = {"l8": 8, "la16": 16, "rgb15": 15, "rgb16": 16, "bgr24": 24, "bgra32": 32}[kind]
        src = rng.integers(0, 256, (w * h, nb)).astype(np.uint8)
        if rle:
            src = np.repeat(src[:: 5], 5, axis=0)[: w * h]                       # runs of five that do not line up with rows
            if len(src) < w * h:
                src = np.concatenate([src, np.repeat(src[-1:], w * h - len(src), axis=0)])
        hdr = header(idlen, 0, itype + (8 if rle else 0), 0, 0, 0, w, h, bpp, desc)
    body = rle_packets(src, rng, src.shape[1]) if rle else src.tobytes()
    if overrun and rle:                                                          # a last packet longer than the image
        body = rle_packets(src[:-3], rng, src.shape[1]) + bytes([0x80 | 9]) + src[-1].tobytes()
        src = np.concatenate([src[:-3], np.repeat(src[-1:], 3, axis=0)])
    return hdr + ident + pal + body, src


def pil_tga(img, rle=False, top_down=False):
    from PIL import Image as PILImage
    b = io.BytesIO()
    im = img if isinstance(img, PILImage.Image) else PILImage.fromarray(img if img.shape[-1] != 1 else img[..., 0])
    im.save(b, "TGA", compression="tga_rle" if rle else None, orientation=1 if top_down else -1)
    return b.getvalue()
