#!/usr/bin/env python
"""Timeline of one gb200_decode_batch_host call (GB200_E2E_TRACE=1) on the PNG (512 x 1080p) and QOIX (256 x 2048^2) workloads."""
import os, sys, time, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import benchlib
from gamut_b200 import codecs
torch.cuda.set_device(0)
for name, n in (("png", 512), ("qoix", 256)):
    wl = benchlib.WORKLOADS[name](0, 1, argparse.Namespace(batch=n, sub_batch=None))
    wl.e2e_n = n
    wl.e2e_setup()
    files = wl.host_files[:n]
    fmt = wl.FORMAT if hasattr(wl, "FORMAT") else 1
    arg = getattr(wl, "E2E_ARG", 0)
    out_bytes = getattr(wl, "out_bytes", None) or wl.out_stride
    for rep in range(3):
        if rep == 2:
            os.environ["GB200_E2E_TRACE"] = "1"
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d = codecs.decode_batch_host(fmt, files, arg, 0, wl.h_out, out_bytes, 0)
        print(name, "total ms", round((time.perf_counter() - t0) * 1e3, 1), file=sys.stderr)
    os.environ.pop("GB200_E2E_TRACE", None)
    wl.e2e_teardown(); wl.release()
