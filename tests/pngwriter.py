"""Minimal PNG writer for tests: any colour type / bit depth, forced or per-row filter types, Adam7,
tRNS/PLTE chunks, several IDAT chunks, zlib level / raw-deflate (CgBI) choice. Pure numpy + zlib."""
from __future__ import annotations

import struct
import zlib

import numpy as np


def _chunk(typ: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + typ + data + struct.pack(">I", zlib.crc32(typ + data) & 0xFFFFFFFF)


def _paeth(a, b, c):
    p = a.astype(np.int32) + b - c
    pa, pb, pc = np.abs(p - a), np.abs(p - b), np.abs(p - c)
    return np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c))


def filter_rows(rows: np.ndarray, bpp: int, filters) -> bytes:
    """rows: (h, rowbytes) uint8 of packed samples. filters: int or sequence of per-row filter types."""
    h, rb = rows.shape
    out = bytearray()
    prev = np.zeros(rb, np.int32)
    for y in range(h):
        f = filters if isinstance(filters, int) else filters[y % len(filters)]
        cur = rows[y].astype(np.int32)
        left = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]]) if rb > bpp else np.zeros(rb, np.int32)[:rb]
        if rb <= bpp:
            left = np.zeros(rb, np.int32)
        upleft = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]]) if rb > bpp else np.zeros(rb, np.int32)
        if f == 0:
            r = cur
        elif f == 1:
            r = cur - left
        elif f == 2:
            r = cur - prev
        elif f == 3:
            r = cur - ((left + prev) >> 1)
        else:
            r = cur - _paeth(left, prev, upleft)
        out.append(f)
        out += (r & 255).astype(np.uint8).tobytes()
        prev = cur
    return bytes(out)


def filter_rows_adaptive(rows: np.ndarray, bpp: int):
    """libpng's default heuristic (PNG spec 12.8): per row, the filter whose residuals, read as signed bytes, have
    the smallest sum of absolute values. Vectorised over the whole image. Returns (raw bytes, per-row filter types)."""
    h, rb = rows.shape
    cur = rows.astype(np.int16)
    left = np.zeros_like(cur)
    left[:, bpp:] = cur[:, :-bpp]
    up = np.zeros_like(cur)
    up[1:] = cur[:-1]
    ul = np.zeros_like(cur)
    ul[1:, bpp:] = cur[:-1, :-bpp]
    p = left + up - ul
    pa, pb, pc = np.abs(p - left), np.abs(p - up), np.abs(p - ul)
    paeth = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, ul))
    cands = [cur, cur - left, cur - up, cur - ((left + up) >> 1), cur - paeth]
    res = [(c & 255).astype(np.uint8) for c in cands]
    cost = np.stack([np.minimum(r.astype(np.int32), 256 - r.astype(np.int32)).sum(axis=1) for r in res])   # (5, h)
    choice = np.argmin(cost, axis=0).astype(np.uint8)
    out = np.empty((h, rb + 1), np.uint8)
    out[:, 0] = choice
    for f in range(5):
        m = choice == f
        out[m, 1:] = res[f][m]
    return out.tobytes(), choice


def pack_samples(img: np.ndarray, depth: int) -> np.ndarray:
    """img: (h, w, c) integer samples (values < 2**depth). Returns (h, rowbytes) uint8, PNG byte order."""
    h, w, c = img.shape
    if depth == 8:
        return img.astype(np.uint8).reshape(h, w * c)
    if depth == 16:
        return img.astype(">u2").view(np.uint8).reshape(h, w * c * 2)
    flat = img.reshape(h, w * c).astype(np.uint8)
    per = 8 // depth
    pad = (-flat.shape[1]) % per
    if pad:
        flat = np.concatenate([flat, np.zeros((h, pad), np.uint8)], axis=1)
    flat = flat.reshape(h, -1, per)
    out = np.zeros(flat.shape[:2], np.uint8)
    for i in range(per):
        out |= (flat[:, :, i] << (8 - depth * (i + 1))).astype(np.uint8)
    return out


ADAM7 = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]


def write_png(img: np.ndarray, color: int, depth: int, filters=0, interlace: bool = False, level: int = 6,
              palette: np.ndarray | None = None, trns: bytes | None = None, idat_split: int = 0,
              cgbi: bool = False, phys: tuple | None = None, iend: bool = True, extra_raw: bytes = b"") -> bytes:
    """img: (h, w, c) samples, c = channels stored in the file (1 for palette)."""
    h, w, c = img.shape
    bpp = max(1, c * depth // 8)
    if not interlace and isinstance(filters, str) and filters == "adaptive":
        raw = filter_rows_adaptive(pack_samples(img, depth), bpp)[0]
    elif not interlace:
        raw = filter_rows(pack_samples(img, depth), bpp, filters)
    else:
        raw = b""
        for (x0, y0, dx, dy) in ADAM7:
            sub = img[y0::dy, x0::dx]
            if sub.shape[0] and sub.shape[1]:
                raw += filter_rows(pack_samples(np.ascontiguousarray(sub), depth), bpp, filters)
    raw += extra_raw
    if cgbi:
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = co.compress(raw) + co.flush()
    else:
        comp = zlib.compress(raw, level)
    out = b"\x89PNG\r\n\x1a\n"
    if cgbi:
        out += _chunk(b"CgBI", b"\x50\x00\x20\x02")
    out += _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color, 0, 0, 1 if interlace else 0))
    if phys:
        out += _chunk(b"pHYs", struct.pack(">IIB", *phys))
    if palette is not None:
        out += _chunk(b"PLTE", palette.astype(np.uint8).tobytes())
    if trns is not None:
        out += _chunk(b"tRNS", trns)
    if idat_split and len(comp) > idat_split:
        for i in range(0, len(comp), idat_split):
            out += _chunk(b"IDAT", comp[i:i + idat_split])
    else:
        out += _chunk(b"IDAT", comp)
    if iend:
        out += _chunk(b"IEND", b"")
    return out
