#!/usr/bin/env python
"""Does running two half-batches concurrently (two host threads, two streams) beat one call? Prints ms per workload."""
import argparse, json, os, sys, threading, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import benchlib
from gamut_b200 import _lib

def run(name, n, parts):
    args = argparse.Namespace(batch=n, sub_batch=n)
    wl = benchlib.WORKLOADS[name](0, 1, args)
    streams = [torch.cuda.Stream() for _ in range(parts)]
    per = (wl.n + parts - 1) // parts
    def work(t):
        a, e = t * per, min(wl.n, (t + 1) * per)
        with torch.cuda.stream(streams[t]):
            b = wl.decode(wl.host_files[a:e], wl.dev_ptrs[a:e], streams[t].cuda_stream)
            assert all(d.status for d in b.images)
            b.free()
    def step():
        ths = [threading.Thread(target=work, args=(t,)) for t in range(parts)]
        for t in ths: t.start()
        for t in ths: t.join()
    for _ in range(3): step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    wl.release()
    return round(float(np.median(ts)), 2)

out = {}
for name, n in (("qoix", 256), ("png", 1024), ("jpeg", 512)):
    out[name] = {p: run(name, n, p) for p in (1, 2, 3)}
    torch.cuda.empty_cache(); _lib.lib().gb200_device_trim()
print(json.dumps(out))
