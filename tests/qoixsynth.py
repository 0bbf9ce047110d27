"""Synthetic streams for the QOIX sub-codecs whose reference encoders are not restated (QOI2AVG 8-bit RGB/RGBA,
QOI-Plane 8-bit L/LA, QOI-10b). Written from the format descriptions (qoi2avg.d:293-303, qoiplane.d:74-90,
qoi10b.d:66-80), independent of the decoders under test:

  * fuzz_*   : random but well-formed opcode streams that use every opcode (GPU vs oracle parity);
  * encode_* : minimal encoders using only the literal opcodes (image -> stream -> decode == image pins the
               container, the literal opcodes and the output stage of the oracle).
"""
import struct

import numpy as np


def header(w, h, version, channels, bitdepth, colorspace=0, compression=0, par=-1.0, dpi=-1.0):
    return b"qoix" + struct.pack(">II", w, h) + bytes([version, channels, bitdepth, colorspace, compression]) + \
        struct.pack(">ff", par, dpi)


class BitWriter:
    def __init__(self):
        self.bits = []

    def put(self, value, n):
        for i in range(n - 1, -1, -1):
            self.bits.append((value >> i) & 1)

    def bytes(self, pad_ones=True):
        b = list(self.bits)
        while len(b) % 8:
            b.append(1 if pad_ones else 0)
        return bytes(int("".join(map(str, b[i:i + 8])), 2) for i in range(0, len(b), 8))


# ---------------------------------------------------------------------------------------------- QOI2AVG
def encode_qoi2avg(img, version=1, **kw):
    h, w, c = img.shape
    out = bytearray(header(w, h, version, c, 8, **kw))
    a_prev = 255
    for px in img.reshape(-1, c):
        r, g, b = int(px[0]), int(px[1]), int(px[2])
        a = int(px[3]) if c == 4 else 255
        if a != a_prev:
            out += bytes([0xFE, r, g, b, a]); a_prev = a
        elif r == g == b:
            out += bytes([0xFC, r])
        else:
            out += bytes([0xFD, r, g, b])
    return bytes(out) + b"\xff" * 4


def fuzz_qoi2avg(w, h, c, seed, version=1):
    rng = np.random.default_rng(seed)
    out = bytearray(header(w, h, version, c, 8))
    n = 0
    while n < w * h:
        k = int(rng.integers(0, 100))
        if c == 4 and rng.integers(0, 6) == 0:
            out.append(0xE8 | int(rng.integers(0, 8)))                       # ADIFF (then another opcode)
        if k < 35:
            out.append(int(rng.integers(0, 0x80))); n += 1                    # LUMA
        elif k < 45:
            out.append(0x80 | int(rng.integers(0, 64))); n += 1               # INDEX
        elif k < 60:
            out += bytes([0xC0 | int(rng.integers(0, 32)), int(rng.integers(0, 256))]); n += 1     # LUMA2
        elif k < 70:
            out += bytes([0xE0 | int(rng.integers(0, 8)), int(rng.integers(0, 256)), int(rng.integers(0, 256))]); n += 1
        elif k < 80:
            r = int(rng.integers(0, 8)); out.append(0xF0 | r); n += r + 1     # RUN
        elif k < 84:
            r = int(rng.integers(0, min(1024, 4 * w))); out += bytes([0xF8 | (r >> 8), r & 255]); n += r + 1   # RUN2
        elif k < 90:
            out += bytes([0xFC, int(rng.integers(0, 256))]); n += 1           # GRAY
        elif k < 96:
            out += bytes([0xFD]) + bytes(rng.integers(0, 256, 3).tolist()); n += 1
        else:
            out += bytes([0xFE]) + bytes(rng.integers(0, 256, 4).tolist()); n += 1
    return bytes(out) + b"\xff" * 4


# ---------------------------------------------------------------------------------------------- QOI-Plane (8 bit)
def _nibbles_to_bytes(nib):
    nib = list(nib)
    if len(nib) % 2:
        nib.append(0xF)
    return bytes((nib[i] << 4) | nib[i + 1] for i in range(0, len(nib), 2))


def encode_qoiplane(img, version=1, **kw):
    h, w, c = img.shape
    nib = []
    a_prev = 255
    for px in img.reshape(-1, c):
        l = int(px[0]); a = int(px[1]) if c == 2 else 255
        if a != a_prev:
            nib += [0xB, 0x0, l >> 4, l & 15, a >> 4, a & 15]; a_prev = a    # LA
        else:
            nib += [0xA, l >> 4, l & 15]                                       # DIRECT
    nib += [0xF] * 9
    return header(w, h, version, c, 8, **kw) + _nibbles_to_bytes(nib) + b"\xff" * 4


def fuzz_qoiplane(w, h, c, seed, version=1):
    rng = np.random.default_rng(seed)
    nib = []
    n = 0
    while n < w * h:
        k = int(rng.integers(0, 100))
        if c == 2 and rng.integers(0, 6) == 0:
            nib += [0xB, int(rng.integers(1, 16))]                             # ADIFF (then another opcode)
        if k < 40:
            nib.append(int(rng.integers(0, 8))); n += 1                        # DIFF1
        elif k < 60:
            nib += [8 | int(rng.integers(0, 2)), int(rng.integers(0, 16))]; n += 1     # DIFF2
        elif k < 70:
            nib += [0xA, int(rng.integers(0, 16)), int(rng.integers(0, 16))]; n += 1   # DIRECT
        elif k < 76:
            nib += [0xB, 0] + rng.integers(0, 16, 4).tolist(); n += 1          # LA
        elif k < 92:
            r = int(rng.integers(0, 3)); nib.append(0xC | r); n += r + 1       # REPEAT1
        else:
            v = int(rng.integers(0, 255)); nib += [0xF, v >> 4, v & 15]; n += v + 4    # REPEAT2 (255 = fill is kept for the end)
    nib += [0xF] * 9
    return header(w, h, version, c, 8) + _nibbles_to_bytes(nib) + b"\xff" * 4


# ---------------------------------------------------------------------------------------------- QOI-10b
def encode_qoi10b(img10, version=1, **kw):
    """img10: (h, w, c) integer array of 10-bit values."""
    h, w, c = img10.shape
    grey = c <= 2
    bw = BitWriter()
    a_prev = 1023
    for px in img10.reshape(-1, c):
        a = int(px[c - 1]) if c in (2, 4) else 1023
        r = int(px[0])
        if a != a_prev:
            bw.put(0xFE, 8); bw.put(r, 10)
            if not grey:
                bw.put(int(px[1]), 10); bw.put(int(px[2]), 10)
            bw.put(a, 10); a_prev = a
        elif grey or (px[0] == px[1] == px[2]):
            bw.put(0xFC, 8); bw.put(r, 10)
        else:
            bw.put(0xFD, 8); bw.put(r, 10); bw.put(int(px[1]), 10); bw.put(int(px[2]), 10)
    for _ in range(5):
        bw.put(0xFF, 8)
    return header(w, h, version, c, 10, **kw) + bw.bytes()


def fuzz_qoi10b(w, h, c, seed, version=1):
    rng = np.random.default_rng(seed)
    grey = c <= 2
    bw = BitWriter()
    n = 0
    R = lambda bits: int(rng.integers(0, 1 << bits))
    while n < w * h:
        k = int(rng.integers(0, 100))
        if c in (2, 4) and rng.integers(0, 6) == 0:
            if rng.integers(0, 2):
                bw.put(0x1D, 5); bw.put(R(5), 5)                               # ADIFF 11101xxxxx
            else:
                bw.put(0x3E, 6); bw.put(R(8), 8)                               # ADIFF2 111110xxxxxxxx
        if k < 25:
            bw.put(0, 1); bw.put(R(5), 5)                                      # LUMA 0ggggg[rrrrbbbb]
            if not grey:
                bw.put(R(8), 8)
            n += 1
        elif k < 45:
            bw.put(2, 2); bw.put(R(4), 4)                                      # LUMA0 10gggg[rrrbbb]
            if not grey:
                bw.put(R(6), 6)
            n += 1
        elif k < 60:
            bw.put(6, 3); bw.put(R(7), 7)                                      # LUMA2
            if not grey:
                bw.put(R(12), 12)
            n += 1
        elif k < 70:
            bw.put(0x1C, 5); bw.put(R(9), 9)                                   # LUMA3
            if not grey:
                bw.put(R(16), 16)
            n += 1
        elif k < 82:
            r = int(rng.integers(0, 7)); bw.put(0xF0 | r, 8); n += r + 1       # RUN
        elif k < 86:
            v = R(8); bw.put(0xF7, 8); bw.put(v, 8); n += v + 8                # long RUN
        elif k < 91:
            bw.put(0xFC, 8); bw.put(R(10), 10); n += 1                         # GRAY
        elif k < 96:
            bw.put(0xFD, 8); bw.put(R(10), 10)
            if not grey:
                bw.put(R(20), 20)
            n += 1
        else:
            bw.put(0xFE, 8); bw.put(R(10), 10)
            if not grey:
                bw.put(R(20), 20)
            bw.put(R(10), 10); n += 1
    for _ in range(5):
        bw.put(0xFF, 8)
    return header(w, h, version, c, 10) + bw.bytes()


def lz4_wrap(stream, oracle):
    """Wrap an uncompressed QOIX stream into the LZ4 container (plugins/qoix.d:251-339)."""
    payload = stream[25:]
    comp = oracle.lz4_compress(payload)
    return stream[:16] + b"\x01" + stream[17:25] + struct.pack(">I", len(payload)) + comp
