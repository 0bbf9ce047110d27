#!/usr/bin/env python
"""Encode throughput of gb200_qoix_encode_batch_device on the config-5 shape (2048x2048 10-bit LA), device resident,
beside the oracle's restatement of qoiplane10_encode on one host thread. Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gamut_b200 import codecs            # noqa: E402
from qoixutil import depth_map_la        # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H = W = 2048
distinct = [depth_map_la(H, W, 100 + i, 2) for i in range(8)]
dev = [torch.from_numpy(distinct[i % 8].view(np.int16)).cuda().clone() for i in range(N)]
cap = codecs.qoix_encode_bound(W, H, 2) + 16
outs = [torch.empty(cap, dtype=torch.uint8, device="cuda") for _ in range(N)]
pin, pout, shapes = [t.data_ptr() for t in dev], [o.data_ptr() for o in outs], [(H, W, 2)] * N
for _ in range(3):
    lens = codecs.qoix_encode_batch_device(pin, shapes, pout)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); lens = codecs.qoix_encode_batch_device(pin, shapes, pout); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = float(np.median(ts))
from oracle import pyoracle             # noqa: E402  (CPU baseline + checker only)
t0 = time.perf_counter(); exp = pyoracle.qoiplane10_encode(distinct[0]); cpu_s = time.perf_counter() - t0
ok = outs[0][:lens[0]].cpu().numpy().tobytes() == exp
print(json.dumps({"workload": f"QOI-Plane10 encode {N} x {W}x{H} la16 (device resident)", "ms": round(ms, 3),
                  "Mpixels_per_s": round(N * H * W / ms / 1e3, 1), "stream_bytes_per_image": int(np.mean(lens)),
                  "cpu_oracle_1thread_Mpixels_per_s": round(H * W / cpu_s / 1e6, 1), "bit_identical_to_oracle": bool(ok)}))
