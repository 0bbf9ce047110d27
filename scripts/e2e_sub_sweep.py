#!/usr/bin/env python
"""End-to-end time of gb200_decode_batch_host against its sub-batch size (GB200_E2E_SUB) on the PNG (512 x 1080p) and
QOIX (256 x 2048^2) workloads; prints ms (median of 3 after a warm-up) per setting."""
import os, sys, time, argparse, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import benchlib
from gamut_b200 import codecs
torch.cuda.set_device(0)
out = {}
for name, n, subs in (("png", 512, (0, 128, 171, 256, 512)), ("qoix", 256, (0, 32, 48, 64, 96, 128))):
    wl = benchlib.WORKLOADS[name](0, 1, argparse.Namespace(batch=n, sub_batch=None))
    wl.e2e_n = n
    wl.e2e_setup()
    files = wl.host_files[:n]
    fmt = wl.FORMAT if hasattr(wl, "FORMAT") else 1
    arg = getattr(wl, "E2E_ARG", 0)
    out_bytes = getattr(wl, "out_bytes", None) or wl.out_stride
    out[name] = {}
    for sub in subs:
        ts = []
        for rep in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            d = codecs.decode_batch_host(fmt, files, arg, 0, wl.h_out, out_bytes, sub)
            ts.append((time.perf_counter() - t0) * 1e3)
        out[name][sub] = round(float(np.median(ts[1:])), 1)
    wl.e2e_teardown(); wl.release()
print(json.dumps(out))
