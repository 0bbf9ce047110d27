"""Synthetic JPEG streams for the tests (PIL / libjpeg-turbo as the encoder only)."""
import io

import numpy as np
from PIL import Image as PILImage


def photo(h, w, c, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, c), np.float32)
    for k in range(c):
        a = np.zeros((h, w), np.float32)
        for _ in range(4):
            fx, fy, ph = rng.uniform(0.01, 0.2), rng.uniform(0.01, 0.2), rng.uniform(0, 6.28)
            a += np.sin(xx * fx + yy * fy + ph) * rng.uniform(0.1, 0.3)
        img[:, :, k] = 0.5 + a * 0.5
    img += rng.normal(0, 0.03, (h, w, c)).astype(np.float32)
    # a few hard edges so that high-frequency coefficients and long codes occur
    img[h // 3: h // 3 + 5, :, :] = 1.0
    img[:, w // 2: w // 2 + 3, :] = 0.0
    return (np.clip(img, 0, 1) * 255 + 0.5).astype(np.uint8)


def encode(img, quality=90, subsampling=0, restart_rows=0, restart_blocks=0, optimize=False, dpi=None, progressive=False):
    b = io.BytesIO()
    im = PILImage.fromarray(img if img.shape[2] != 1 else img[:, :, 0])
    kw = dict(quality=quality, optimize=optimize)
    if progressive:
        kw["progressive"] = True
    if img.shape[2] == 3:
        kw["subsampling"] = subsampling
    if restart_rows:
        kw["restart_marker_rows"] = restart_rows
    if restart_blocks:
        kw["restart_marker_blocks"] = restart_blocks
    if dpi:
        kw["dpi"] = dpi
    im.save(b, "JPEG", **kw)
    return b.getvalue()
