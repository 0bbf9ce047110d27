#!/usr/bin/env python
"""Where does the JPEG end-to-end time go? Times gb200_decode_batch_host on 256 4K files for several sub-batch sizes,
one host-file decode call per sub-batch size (no download), and a plain pinned D2H of the same bytes."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import benchlib
from gamut_b200 import codecs


class A: batch = 256; sub_batch = None


def main():
    torch.cuda.set_device(0)
    wl = benchlib.JpegWorkload(0, 1, A())
    wl.e2e_n = 256
    wl.e2e_setup()
    out = {}
    files = wl.host_files[:256]
    for sub in (16, 32, 64, 128, 256):
        ts = []
        for rep in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            d = codecs.decode_batch_host(wl.FORMAT, files, wl.E2E_ARG, 0, wl.h_out, wl.out_bytes, sub)
            ts.append((time.perf_counter() - t0) * 1e3)
            assert all(x.status for x in d)
        out[f"batch_host_sub{sub}_ms"] = [round(t, 1) for t in ts[1:]]
    for m in (16, 32, 64, 128):
        ts = []
        for rep in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            b = codecs.jpeg_decode_batch(files[:m], wl.E2E_ARG)
            ts.append((time.perf_counter() - t0) * 1e3)
            ph, hp = b.timing()
            b.free()
        out[f"decode_call_{m}_images_ms"] = [round(t, 2) for t in ts[1:]]
        out[f"decode_call_{m}_phases"] = [round(x, 2) for x in ph[:8]] + [round(hp, 2)]
    n = 256 * wl.out_bytes
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize()
        out.setdefault("plain_d2h_ms", []).append(round((time.perf_counter() - t0) * 1e3, 1))
    # per-image copies like the library issues them
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(256):
            h[i * wl.out_bytes:(i + 1) * wl.out_bytes].copy_(d[i * wl.out_bytes:(i + 1) * wl.out_bytes], non_blocking=True)
        torch.cuda.synchronize()
        out["per_image_d2h_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
    print(json.dumps(out))


main()
