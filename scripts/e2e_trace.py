#!/usr/bin/env python
"""Timeline of one gb200_decode_batch_host call (GB200_E2E_TRACE=1): 256 4K JPEG files, sub-batch from argv."""
import os, sys, time
os.environ["GB200_E2E_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import benchlib
from gamut_b200 import codecs
class A: batch = 256; sub_batch = None
torch.cuda.set_device(0)
wl = benchlib.JpegWorkload(0, 1, A())
wl.e2e_n = 256
wl.e2e_setup()
files = wl.host_files[:256]
sub = int(sys.argv[1]) if len(sys.argv) > 1 else 32
os.environ.pop("GB200_E2E_TRACE")
for rep in range(2):
    if rep == 1:
        os.environ["GB200_E2E_TRACE"] = "1"
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d = codecs.decode_batch_host(wl.FORMAT, files, wl.E2E_ARG, 0, wl.h_out, wl.out_bytes, sub)
    print("total ms", round((time.perf_counter() - t0) * 1e3, 1), file=sys.stderr)
