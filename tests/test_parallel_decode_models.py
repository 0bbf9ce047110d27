"""CPU models of the ideas behind the parallel device decoders (no GPU, no product code on the path): they pin the
*properties* the CUDA kernels rely on, with independent libraries (zlib via ctypes, liblz4) as the judges.

 * paeth: the 26-instruction packed-byte Paeth of csrc/png_kernels.cu is a re-derivation (two compares instead of
   five); checked against the PNG/stb formula (stbdec.d:1390-1401) for all 2^24 (a, b, c);
 * deflate: every dynamic-Huffman block of a zlib stream is found by the brute-force header search of
   csrc/inflate_par.cuh (BTYPE = 2, HLIT/HDIST <= 29, complete code-length code, complete literal/length and distance
   codes), judged by zlib's own block boundaries (inflate with Z_BLOCK);
 * lz4: the token chain of an LZ4 block synchronises itself: a walk started at an arbitrary byte meets the true chain
   (csrc/qoix.cu lz4_spec / lz4_merge kernels)."""
import ctypes as C
import ctypes.util
import zlib

import numpy as np
import pytest


def test_paeth_two_compare_identity():
    a, b, c = np.meshgrid(np.arange(256, dtype=np.int32), np.arange(256, dtype=np.int32), np.arange(256, dtype=np.int32), indexing="ij")
    a, b, c = a.ravel(), b.ravel(), c.ravel()
    p = a + b - c
    pa, pb, pc = np.abs(p - a), np.abs(p - b), np.abs(p - c)
    ref = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c))
    PA, PB, T = np.abs(b - c), np.abs(a - c), np.abs(a - b)
    U = np.abs(PA - PB)
    le = PA <= PB
    q = np.where(le, PA, PB)
    V = np.where(T == U, 255, U)            # same sign of a-c and b-c  <=>  |a-b| == ||b-c| - |a-c||
    got = np.where(q <= V, np.where(le, a, b), c)
    assert np.array_equal(got, ref)


# ---- deflate block boundaries from zlib itself ---------------------------------------------------------------
class _ZStream(C.Structure):
    _fields_ = [("next_in", C.c_void_p), ("avail_in", C.c_uint), ("total_in", C.c_ulong), ("next_out", C.c_void_p),
                ("avail_out", C.c_uint), ("total_out", C.c_ulong), ("msg", C.c_char_p), ("state", C.c_void_p),
                ("zalloc", C.c_void_p), ("zfree", C.c_void_p), ("opaque", C.c_void_p), ("data_type", C.c_int),
                ("adler", C.c_ulong), ("reserved", C.c_ulong)]


def _block_starts(z: bytes):
    """Bit positions at which zlib's inflate reports a block boundary (Z_BLOCK), i.e. where block headers start."""
    name = ctypes.util.find_library("z")
    if not name:
        pytest.skip("libz not found")
    L = C.CDLL(name)
    s = _ZStream()
    L.zlibVersion.restype = C.c_char_p
    assert L.inflateInit_(C.byref(s), L.zlibVersion(), C.sizeof(_ZStream)) == 0
    src = C.create_string_buffer(z, len(z))
    out = C.create_string_buffer(1 << 20)
    s.next_in = C.cast(src, C.c_void_p); s.avail_in = len(z)
    starts, last = [], []
    while True:
        s.next_out = C.cast(out, C.c_void_p); s.avail_out = len(out)
        r = L.inflate(C.byref(s), 5)                      # Z_BLOCK
        if s.data_type & 128:
            starts.append(int(s.total_in) * 8 - (s.data_type & 63))
            last.append(bool(s.data_type & 64))
        if r == 1 or r < 0:
            break
    L.inflateEnd(C.byref(s))
    assert r == 1
    # the first report is the end of the zlib header = start of block 0; the final one is the end of the stream
    return starts[:-1]


_ORDER = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]


def _header_ok(bits: np.ndarray, p: int) -> bool:
    """The find + verify filters of inflate_par.cuh at bit position p (bits: one uint8 per stream bit)."""
    def get(q, n):
        v = 0
        for i in range(n):
            v |= int(bits[q + i]) << i
        return v
    if p + 17 + 57 > len(bits) or get(p + 1, 2) != 2:
        return False
    hlit, hdist, hclen = get(p + 3, 5), get(p + 8, 5), get(p + 13, 4)
    if hlit > 29 or hdist > 29:
        return False
    cl = [0] * 19
    q = p + 17
    for i in range(hclen + 4):
        cl[_ORDER[i]] = get(q, 3); q += 3
    if sum(128 >> l for l in cl if l) != 128:
        return False
    # canonical code of the code-length code
    codes, code = {}, 0
    for l in range(1, 8):
        for sym in range(19):
            if cl[sym] == l:
                codes[(l, code)] = sym; code += 1
        code <<= 1
    total, nlit = hlit + 257 + hdist + 1, hlit + 257
    lens, prev = [], 0
    while len(lens) < total:
        code = 0; sym = None
        for l in range(1, 8):
            if q >= len(bits):
                return False
            code = (code << 1) | int(bits[q]); q += 1
            if (l, code) in codes:
                sym = codes[(l, code)]; break
        if sym is None:
            return False
        if sym < 16:
            lens.append(sym); prev = sym; continue
        if sym == 16:
            if not lens:
                return False
            rep, val = 3 + get(q, 2), prev; q += 2
        elif sym == 17:
            rep, val = 3 + get(q, 3), 0; q += 3
        else:
            rep, val = 11 + get(q, 7), 0; q += 7
        if len(lens) + rep > total:
            return False
        lens += [val] * rep; prev = val
    lit, dist = lens[:nlit], lens[nlit:]
    if lit[256] == 0 or sum(32768 >> l for l in lit if l) != 32768:
        return False
    nd = sum(1 for l in dist if l)
    return sum(32768 >> l for l in dist if l) == 32768 or nd <= 1


def test_deflate_dynamic_headers_are_found():
    rng = np.random.default_rng(3)
    x = np.arange(400_000)
    data = ((128 + 70 * np.sin(x / 29.0) + rng.normal(0, 6, x.size)).clip(0, 255)).astype(np.uint8).tobytes()
    data += bytes(20_000) + rng.integers(0, 256, 30_000, dtype=np.uint8).tobytes()      # + a run, + a stored block
    z = zlib.compress(data, 6)
    bits = np.unpackbits(np.frombuffer(z + bytes(16), np.uint8), bitorder="little")
    starts = _block_starts(z)
    assert len(starts) >= 10
    dyn = [p for p in starts if int(bits[p + 1]) | (int(bits[p + 2]) << 1) == 2]
    assert len(dyn) >= len(starts) - 4                     # level 6 emits dynamic blocks except for the stored part
    for p in dyn:
        assert _header_ok(bits, p), p
    # and the filter is selective: none of a few thousand other positions passes
    other = [int(q) for q in rng.integers(16, len(z) * 8 - 4000, 3000) if int(q) not in set(starts)]
    assert sum(_header_ok(bits, q) for q in other) == 0


# ---- LZ4 token chain self-synchronisation --------------------------------------------------------------------
def _lz4_step(b, x):
    tok = b[x]; x += 1
    L = tok >> 4
    if L == 15:
        while True:
            s = b[x]; x += 1; L += s
            if s != 255:
                break
    x += L
    if x >= len(b):
        return len(b)
    x += 2
    M = tok & 15
    if M == 15:
        while True:
            s = b[x]; x += 1
            if s != 255:
                break
    return x


def test_lz4_token_chain_self_synchronises():
    from qoixutil import liblz4
    L = liblz4()
    if L is None:
        pytest.skip("liblz4 not found")
    rng = np.random.default_rng(9)
    words = [bytes(rng.integers(0, 256, int(rng.integers(2, 12)), dtype=np.uint8)) for _ in range(500)]
    src = b"".join(words[int(i)] for i in rng.integers(0, 500, 60_000)) + rng.integers(0, 256, 5000, dtype=np.uint8).tobytes()
    cap = L.LZ4_compressBound(len(src))
    dst = C.create_string_buffer(cap)
    n = L.LZ4_compress_default(src, dst, len(src), cap)
    b = dst.raw[:n]
    true_pos, x = set(), 0
    while x < len(b):
        true_pos.add(x); x = _lz4_step(b, x)
    met = 0
    starts = [int(q) for q in rng.integers(1, len(b) - 4096, 200)]
    for q in starts:
        x, steps = q, 0
        while x < len(b) and x not in true_pos and x < q + 4096:
            try:
                x = _lz4_step(b, x)
            except IndexError:
                break
            steps += 1
        met += x in true_pos and x < len(b)
    assert met >= 0.97 * len(starts)        # the rare miss is what the merge kernel's true-walk fallback is for


# ---- QOI-Plane10 encoder: every code from four input pixels + a prefix maximum + a prefix sum ------------------
def _p10_encode_model(img):
    """numpy model of csrc/qoix_encode.cu: per-pixel codes computed independently (predictor of the ORIGINAL
    neighbours), run positions from the distance to the last pixel that differs from its predecessor (a prefix
    maximum), bit positions from the prefix sum of the code lengths. Returns the payload bits as a string."""
    h, w, c = img.shape
    v = (img.astype(np.int64) >> 6)
    l = v[:, :, 0].reshape(-1)
    a = v[:, :, 1].reshape(-1) if c == 2 else np.full(h * w, 1023, np.int64)
    n = h * w
    pl = np.concatenate(([0], l[:-1])); pa = np.concatenate(([1023], a[:-1]))          # previous pixel in raster order
    idx = np.arange(n); y = idx // w; x = idx % w
    up = np.where(y > 0, l[np.maximum(idx - w, 0)], 0)
    upleft = np.where((y > 0) & (x > 0), l[np.maximum(idx - w - 1, 0)], 0)
    mx, mn = np.maximum(pl, up), np.minimum(pl, up)
    med = np.where(upleft >= mx, mn, np.where(upleft <= mn, mx, np.clip(pl + up - upleft, 0, 1023)))
    pred = np.where(y == 0, pl, np.where(x == 0, up, med))
    eq = (l == pl) & (a == pa)
    last_ne = np.maximum.accumulate(np.where(eq, -1, idx))                             # the prefix maximum
    r = (idx - (last_ne + 1)) & 255
    nxt_eq = np.concatenate((eq[1:], [False]))
    run_end = eq & ((r == 255) | (idx == n - 1) | ~nxt_eq)
    vg = (l - pred) & 1023
    va = (a - pa) & 1023
    small = (vg < 4) | (vg >= 1020)
    codes = []
    for i in range(n):                      # the emit step, pixel by pixel (each entry depends on pixel i only)
        if eq[i]:
            if not run_end[i]:
                codes.append("")
            elif r[i] == 0 and small[i]:
                codes.append(format(int(vg[i]) & 7, "04b"))
            elif r[i] < 7:
                codes.append(format(0x30 | int(r[i]), "06b"))
            else:
                codes.append(format(0x37, "06b") + format(int(r[i]) - 7, "08b"))
            continue
        s = ""
        if va[i]:
            if va[i] < 32 or va[i] >= 992:
                s = format((0x3e << 6) | (int(va[i]) & 0x3f), "012b")
            else:
                codes.append(format(0xfe, "08b") + format(int(l[i]), "010b") + format(int(a[i]), "010b"))
                continue
        g = int(vg[i])
        if small[i]:
            s += format(g & 7, "04b")
        elif g < 32 or g >= 992:
            s += format(0x80 | (g & 0x3f), "08b")
        elif g < 64 or g >= 960:
            s += format((0x1e << 7) | (g & 0x7f), "012b")
        else:
            s += format((0xe << 10) | g, "014b")
        codes.append(s)
    lens = np.array([len(s) for s in codes])
    offs = np.concatenate(([0], np.cumsum(lens)))                                       # the prefix sum
    bits = ["0"] * int(offs[-1])
    for i, s in enumerate(codes):           # any order: every code knows its place
        bits[int(offs[i]):int(offs[i + 1])] = s
    payload = "".join(bits) + "1" * 40
    return payload + "1" * (-len(payload) % 8)


def test_qoiplane10_encoder_parallel_formulation_equals_the_serial_encoder():
    from oracle import pyoracle
    from qoixutil import depth_map_la
    rng = np.random.default_rng(4)
    imgs = [depth_map_la(19, 37, 1, 2), depth_map_la(8, 300, 2, 1)]
    flat = np.zeros((9, 70, 2), np.int64) + np.array([517, 1023])
    flat[4, 10:] = [518, 1023]
    imgs.append(((flat << 6) | (flat >> 4)).astype(np.uint16))                          # runs across rows, cut at 256
    noise = rng.integers(0, 1024, (11, 13, 2))
    imgs.append(((noise << 6) | (noise >> 4)).astype(np.uint16))
    for img in imgs:
        exp = pyoracle.qoiplane10_encode(img)
        bits = "".join(format(b, "08b") for b in exp[25:])
        assert _p10_encode_model(img) == bits


# ---- QOI encoder: the index hit of a pixel is a "previous occurrence with the same hash" query ------------------
def _qoi_encode_model(img):
    """numpy model of csrc/qoi_encode.cu. qoi_encode (qoi.d:295-426) keeps index[hash] = the latest pixel with that
    hash that was not a run pixel, so "INDEX" for pixel i is: the latest earlier non-run pixel with i's hash has i's
    value (or there is none and the pixel is all zero, the index's initial content). Everything else depends on pixels
    i-1 and i only; runs are cut every 62 pixels from the start of the maximal sequence of equal pixels."""
    h, w, c = img.shape
    px = np.zeros((h * w, 4), np.int64); px[:, 3] = 255
    px[:, :c] = img.reshape(-1, c)
    v = px[:, 0] | (px[:, 1] << 8) | (px[:, 2] << 16) | (px[:, 3] << 24)
    n = h * w
    pv = np.concatenate(([255 << 24], v[:-1]))
    pp = np.concatenate(([[0, 0, 0, 255]], px[:-1]))
    idx = np.arange(n)
    eq = v == pv
    last_ne = np.maximum.accumulate(np.where(eq, -1, idx))
    r = (idx - (last_ne + 1)) % 62
    run_end = eq & ((r == 61) | (idx == n - 1) | ~np.concatenate((eq[1:], [False])))
    hsh = (px[:, 0] * 3 + px[:, 1] * 5 + px[:, 2] * 7 + px[:, 3] * 11) % 64
    # previous non-run pixel with the same hash (per-bucket prefix maximum over the non-run pixels)
    prev_same = np.full(n, -1)
    for b in range(64):
        sel = np.flatnonzero(~eq & (hsh == b))
        prev_same[sel[1:]] = sel[:-1]
    hit = ~eq & np.where(prev_same >= 0, v[np.maximum(prev_same, 0)] == v, v == 0)
    s8 = lambda x: ((x + 128) & 255) - 128
    vr, vg, vb = s8(px[:, 0] - pp[:, 0]), s8(px[:, 1] - pp[:, 1]), s8(px[:, 2] - pp[:, 2])
    vgr, vgb = s8(vr - vg), s8(vb - vg)
    out = []
    for i in range(n):
        if eq[i]:
            out.append(bytes([0xc0 | int(r[i])]) if run_end[i] else b"")
        elif hit[i]:
            out.append(bytes([int(hsh[i])]))
        elif px[i, 3] != pp[i, 3]:
            out.append(bytes([0xff, *px[i].tolist()]))
        elif -3 < vr[i] < 2 and -3 < vg[i] < 2 and -3 < vb[i] < 2:
            out.append(bytes([0x40 | (int(vr[i]) + 2) << 4 | (int(vg[i]) + 2) << 2 | (int(vb[i]) + 2)]))
        elif -9 < vgr[i] < 8 and -33 < vg[i] < 32 and -9 < vgb[i] < 8:
            out.append(bytes([0x80 | (int(vg[i]) + 32), (int(vgr[i]) + 8) << 4 | (int(vgb[i]) + 8)]))
        else:
            out.append(bytes([0xfe, *px[i, :3].tolist()]))
    return b"".join(out)


def test_qoi_encoder_parallel_formulation_equals_the_serial_encoder():
    from oracle import pyoracle
    from qoixutil import qoi_test_image
    rng = np.random.default_rng(2)
    imgs = [qoi_test_image(40, 50, 4, 1), qoi_test_image(33, 70, 3, 2), np.zeros((6, 40, 4), np.uint8),
            np.zeros((3, 70, 3), np.uint8), rng.integers(0, 4, (30, 30, 4)).astype(np.uint8) * 60]
    start255 = np.zeros((4, 30, 4), np.uint8); start255[..., 3] = 255; start255[2:, :, 0] = 9      # starts inside the initial run
    imgs.append(start255)
    for img in imgs:
        exp = pyoracle.qoi_encode(img)
        assert _qoi_encode_model(img) == exp[14:-8]
