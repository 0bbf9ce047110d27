"""One call of each of the kernels added at the end of round 2 (QOI / QOI-Plane / QOI-Plane10 encoders, TGA decoder and
encoder) on 16 images each, meant to run under the ncu launch-list pass:

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_new_kernels.csv \
        python scripts/new_kernels_launches.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
from gamut_b200 import codecs
from qoixutil import depth_map_la, qoi_test_image
from tgautil import pil_tga

N = 16
rng = np.random.default_rng(0)


def batch(imgs, bound):
    dev = [torch.from_numpy(i.view(np.int16) if i.itemsize == 2 else i).cuda() for i in imgs]
    outs = [torch.empty(bound, dtype=torch.uint8, device="cuda") for _ in imgs]
    return dev, outs, [t.data_ptr() for t in dev], [o.data_ptr() for o in outs], [i.shape for i in imgs]


q = [qoi_test_image(1080, 1920, 4, 100 + (k & 3)) for k in range(N)]
dev, outs, p, o, sh = batch(q, codecs.qoi_encode_bound(1920, 1080, 4) + 16)
for _ in range(2):
    codecs.qoi_encode_batch_device(p, sh, o)
la = [depth_map_la(2048, 2048, 200 + (k & 1), 2) for k in range(N)]
dev, outs, p, o, sh = batch(la, codecs.qoix_encode_bound(2048, 2048, 2) + 16)
for _ in range(2):
    codecs.qoix_encode_batch_device(p, sh, o)
la8 = [(i >> 8).astype(np.uint8) for i in la]
dev, outs, p, o, sh = batch(la8, codecs.qoix_encode_bound(2048, 2048, 2) + 16)
for _ in range(2):
    codecs.qoix_encode_batch_device(p, sh, o, bitdepths=[8] * N)
img = np.zeros((1080, 1920, 4), np.uint8)
img[..., :3] = (np.linspace(0, 255, 1920)[None, :, None] + rng.integers(0, 3, (1080, 1920, 1))).astype(np.uint8)
img[200:700, 300:1500, :3] = rng.integers(0, 256, (500, 1200, 3))
img[..., 3] = 255
for rle in (False, True):
    files = [pil_tga(img, rle)] * N
    fd = [torch.frombuffer(bytearray(f), dtype=torch.uint8).cuda() for f in files]
    for _ in range(2):
        codecs.tga_decode_batch(files, files_dev=[t.data_ptr() for t in fd]).free()
dev, outs, p, o, sh = batch([img] * N, codecs.tga_encode_bound(1920, 1080, 4))
for _ in range(2):
    codecs.tga_encode_batch_device(p, sh, o)
torch.cuda.synchronize()
print("done")
