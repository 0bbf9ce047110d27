"""Synthetic QOIX / QOI inputs for the tests."""
import ctypes as C
import ctypes.util
import io

import numpy as np


def depth_map_la(h, w, seed, channels=2):
    """Depth-map-like 10-bit luma (smooth + sparse edges + noise), alpha mostly 1023 with soft-edged regions.
    Returns (h, w, c) uint16 with 10-bit values expanded as (v<<6)|(v>>4) (lossless in QOI-Plane10)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    l = 500 + 300 * np.sin(xx / 37.0 + seed) * np.cos(yy / 53.0) + rng.normal(0, 1.2, (h, w))
    l[h // 4: h // 2, w // 3: w // 2] += 180                      # an object edge
    l[:, : w // 8] = 100                                          # flat region -> runs
    l = np.clip(l, 0, 1023).astype(np.int64)
    a = np.full((h, w), 1023, np.int64)
    r = np.hypot(xx - w * 0.7, yy - h * 0.6)
    a = np.where(r < min(h, w) * 0.2, np.clip((r / (min(h, w) * 0.2)) * 1023, 0, 1023).astype(np.int64), a)
    a[max(0, h - 3):, :] = rng.integers(0, 1024, (min(3, h), w))  # big alpha jumps -> LA opcodes
    v = np.stack([l, a], axis=2)[:, :, :channels]
    return ((v << 6) | (v >> 4)).astype(np.uint16)


def liblz4():
    for name in ("liblz4.so.1", ctypes.util.find_library("lz4")):
        if not name:
            continue
        try:
            L = C.CDLL(name)
            L.LZ4_compress_default.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
            L.LZ4_decompress_safe.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
            L.LZ4_compressBound.argtypes = [C.c_int]
            return L
        except OSError:
            continue
    return None


def qoi_bytes(img):
    from PIL import Image as PILImage
    b = io.BytesIO()
    PILImage.fromarray(img).save(b, "QOI")
    return b.getvalue()


def qoi_test_image(h, w, c, seed):
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w, c), np.uint8)
    img[..., :3] = np.linspace(0, 255, w)[None, :, None].astype(np.uint8)      # smooth gradient: DIFF/LUMA
    img[..., 1] += rng.integers(0, 3, (h, w)).astype(np.uint8)                 # low-amplitude noise
    img[h // 4: h // 2, w // 4: w // 2, :3] = [200, 10, 10]                    # flat rectangle: RUN / INDEX
    img[h // 2:, : w // 3, :3] = rng.integers(0, 256, (h - h // 2, w // 3, 3)) # noise: RGB ops
    if c == 4:
        img[..., 3] = np.linspace(0, 255, h)[:, None].astype(np.uint8)         # alpha ramp: RGBA ops
        img[: h // 8, :, 3] = 255
    return img
