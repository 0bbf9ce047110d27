"""CPU restatements (oracle/, one thread) of the encoders and of the TGA decoder on the images scripts/encoders_gpu_pass.py
and scripts/tga_gpu_pass.py use: median of 5 calls after one warm-up. Runs anywhere (no GPU); the host it ran on is
recorded. Writes profiles/r2_encoders_cpu_baseline.json."""
import json
import os
import platform
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
from oracle import pyoracle
from qoixutil import depth_map_la, qoi_test_image
from tgautil import pil_tga


def med(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return sorted(ts)[len(ts) // 2]


rgba = qoi_test_image(1080, 1920, 4, 100)
rgba16 = (qoi_test_image(1080, 1920, 4, 400).astype(np.uint16) * 257) & 0xffc0
la8 = (depth_map_la(2048, 2048, 200, 2) >> 8).astype(np.uint8)
rng = np.random.default_rng(0)
photo = np.zeros((1080, 1920, 4), np.uint8)
photo[..., :3] = (np.linspace(0, 255, 1920)[None, :, None] + rng.integers(0, 3, (1080, 1920, 1))).astype(np.uint8)
photo[200:700, 300:1500, :3] = rng.integers(0, 256, (500, 1200, 3))
photo[..., 3] = 255
tga_raw, tga_rle = pil_tga(photo, False), pil_tga(photo, True)
cases = [("qoi_encode", rgba.shape, lambda: pyoracle.qoi_encode(rgba)), ("qoi2avg_encode", rgba.shape, lambda: pyoracle.qoi2avg_encode(rgba)),
         ("qoi10b_encode", rgba16.shape, lambda: pyoracle.qoi10b_encode(rgba16)), ("qoiplane_encode", la8.shape, lambda: pyoracle.qoiplane_encode(la8)),
         ("tga_encode", photo.shape, lambda: pyoracle.tga_encode(photo)), ("bmp_encode", rgba.shape, lambda: pyoracle.bmp_encode(rgba)),
         ("tga_decode_raw", photo.shape, lambda: pyoracle.tga_load(tga_raw)), ("tga_decode_rle", photo.shape, lambda: pyoracle.tga_load(tga_rle))]
out = {"host": platform.processor() or platform.machine(), "cores_used": 1, "kind": "port (oracle C restatement, gcc -O2), includes the ctypes call and the copy of the result"}
for name, shape, fn in cases:
    s = med(fn)
    out[name] = {"image": list(shape), "ms": s * 1e3, "Mpx_per_s": shape[0] * shape[1] / s / 1e6}
    print(name, out[name])
with open(os.path.join(ROOT, "profiles", "r2_encoders_cpu_baseline.json"), "w") as f:
    json.dump(out, f, indent=1)
