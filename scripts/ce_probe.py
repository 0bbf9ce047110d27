#!/usr/bin/env python
"""Do a host-to-device copy and a kernel on one stream overlap a long device-to-host copy on another stream?"""
import json, time, torch
torch.cuda.set_device(0)
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda"); h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
m = 64 << 20
d2 = torch.empty(m, dtype=torch.uint8, device="cuda"); h2 = torch.empty(m, dtype=torch.uint8, pin_memory=True)
hp = torch.empty(1 << 16, dtype=torch.uint8)          # pageable
dp = torch.empty(1 << 16, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
out = {}
def ev(): return torch.cuda.Event(enable_timing=True)
for name in ("pinned_h2d_64MB", "pageable_h2d_64KB", "pageable_d2h_64KB", "kernel_only", "mapped_read_kernel_64MB"):
    for rep in range(2):
        torch.cuda.synchronize()
        a0, a1, b0, b1 = ev(), ev(), ev(), ev()
        with torch.cuda.stream(s1):
            a0.record(s1); h.copy_(d, non_blocking=True); a1.record(s1)
        t0 = time.perf_counter()
        with torch.cuda.stream(s2):
            b0.record(s2)
            if name == "pinned_h2d_64MB": d2.copy_(h2, non_blocking=True)
            elif name == "pageable_h2d_64KB": dp.copy_(hp, non_blocking=True)
            elif name == "pageable_d2h_64KB": hp.copy_(dp, non_blocking=True)
            elif name == "kernel_only": d2.add_(1)
            else:
                # a kernel that reads pinned host memory directly (UVA zero-copy): torch cannot wrap a host pointer as
                # a cuda tensor, so use cudaHostGetDevicePointer through ctypes-free trick: skip if unsupported
                try:
                    import ctypes
                    rt = ctypes.CDLL("libcudart.so.12")
                    dptr = ctypes.c_void_p()
                    rt.cudaHostGetDevicePointer(ctypes.byref(dptr), ctypes.c_void_p(h2.data_ptr()), 0)
                    rt.cudaMemcpyAsync(ctypes.c_void_p(d2.data_ptr()), dptr, ctypes.c_size_t(m), 3, ctypes.c_void_p(s2.cuda_stream))  # D2D kind on a mapped host pointer: runs as a copy, not a kernel
                except Exception as e:
                    out[name] = str(e)
            b1.record(s2)
        host_ms = (time.perf_counter() - t0) * 1e3
        torch.cuda.synchronize()
    out[name] = {"d2h_1GiB_ms": round(a0.elapsed_time(a1), 2), "other_start_after_d2h_start_ms": round(a0.elapsed_time(b0), 2),
                 "other_end_after_d2h_start_ms": round(a0.elapsed_time(b1), 2), "host_call_ms": round(host_ms, 2)}
print(json.dumps(out))
