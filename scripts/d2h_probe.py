#!/usr/bin/env python
"""Device-to-host bandwidth per rank and in aggregate: copy engine (cudaMemcpyAsync to pinned memory) against stores
from the SMs into mapped pinned memory (gb200_download_by_kernel). Run under torchrun for N > 1."""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from gamut_b200 import _lib
L = _lib.lib()
L.gb200_host_alloc.restype = C.c_void_p; L.gb200_host_alloc.argtypes = [C.c_size_t]
L.gb200_download_by_kernel.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda").random_(0, 255)
hp = L.gb200_host_alloc(n)
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
s = torch.cuda.Stream()
out = {}
def bar():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
for name in ("copy_engine", "sm_stores", "both_halves"):
    best = 1e9
    for rep in range(3):
        bar()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            a.record(s)
            if name == "copy_engine": h.copy_(d, non_blocking=True)
            elif name == "sm_stores": assert L.gb200_download_by_kernel(hp, d.data_ptr(), n, s.cuda_stream)
            else:
                assert L.gb200_download_by_kernel(hp, d.data_ptr(), n // 2, s.cuda_stream)
                h[n // 2:].copy_(d[n // 2:], non_blocking=True)      # same stream: serial; measures nothing new but checks mixing
            b.record(s)
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    t = torch.tensor([best], device="cuda", dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[name] = {"ms_max_over_ranks": round(float(t.item()), 2), "aggregate_GBps": round(n * world / (float(t.item()) * 1e-3) / 1e9, 1)}
# check the kernel path delivered the bytes
import numpy as np
got = np.ctypeslib.as_array((C.c_uint8 * 4096).from_address(hp))
out["sm_stores_correct"] = bool((torch.from_numpy(got.copy()) == d[:4096].cpu()).all())
if rank == 0:
    print(json.dumps({"world": world, **out}))
