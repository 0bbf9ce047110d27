#!/usr/bin/env python
"""bench.py -- hot-path throughput on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload convert|png|jpeg|qoix] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input. Default workload is
BASELINE.json configs[1]: PixelType convert rgba8<->rgbaf32 on one 8192x8192 image (forward +
reverse = 2 launches, 134.2 Mpixels per step). One process per GPU (torchrun for N>1), batch units
sharded across ranks with no data-path collective ("weak" scaling: every rank converts its own image).

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region (NVML, in-process, every ~2 ms;
    same fields as the nvidia-smi query of B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.mx = None
        self.stop_flag = False
        self.t = None
        self.err = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except Exception:
                    pass
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                     "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

            def run():
                while not self.stop_flag:
                    try:
                        self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        try:
                            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                        except Exception:
                            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for k, bit in names.items():
                            if r & bit:
                                self.reasons.add(k)
                    except Exception as e:  # noqa: BLE001
                        self.err = str(e)
                        break
                    time.sleep(0.002)

            self.t = threading.Thread(target=run, daemon=True)
            self.t.start()
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    def stop(self):
        self.stop_flag = True
        if self.t:
            self.t.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.mx, "samples": 0, "reasons": ["nvml unavailable: %s" % self.err]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.mx, "samples": len(self.samples),
                "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------------
# Workload: config 2 -- PixelType convert rgba8 <-> rgbaf32, 8192x8192
class ConvertWorkload:
    name = "PixelType convert rgba8<->rgbaf32 8192x8192 (BASELINE configs[1])"
    dtype = "f32"
    W = H = 8192
    bytes_per_px = 20            # 4 B rgba8 + 16 B rgbaf32, per direction
    launches_per_step = 2

    def __init__(self, rank: int):
        import torch
        from gamut_b200 import _lib
        self.torch = torch
        self.L = _lib.lib()
        W, H = self.W, self.H
        g = torch.Generator(device="cuda").manual_seed(1 + rank)
        self.u8 = torch.randint(0, 256, (H, W, 4), dtype=torch.uint8, device="cuda", generator=g)
        self.f32 = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
        self.f32_in = torch.rand((H, W, 4), dtype=torch.float32, device="cuda", generator=g)   # seed-2 style case
        self.u8_out = torch.empty_like(self.u8)
        self.px_per_step = 2 * W * H
        self.kernel_ms = {"convert_direct<rgba8,rgbaf32>": [], "convert_direct<rgbaf32,rgba8>": []}
        self.host = None

    def step(self, stream, events=None):
        from gamut_b200.types import PixelType as PT
        W, H, L = self.W, self.H, self.L
        st = stream.cuda_stream
        if events is not None:
            events[0].record(stream)
        ok1 = L.gb200_scanlines_convert_device(PT.rgba8, self.u8.data_ptr(), W * 4, PT.rgbaf32, self.f32.data_ptr(), W * 16, W, H, st)
        if events is not None:
            events[1].record(stream)
        ok2 = L.gb200_scanlines_convert_device(PT.rgbaf32, self.f32_in.data_ptr(), W * 16, PT.rgba8, self.u8_out.data_ptr(), W * 4, W, H, st)
        if events is not None:
            events[2].record(stream)
        if not (ok1 and ok2):
            raise RuntimeError(self.L.gb200_last_error().decode())

    def collect(self, events):
        self.kernel_ms["convert_direct<rgba8,rgbaf32>"].append(events[0].elapsed_time(events[1]))
        self.kernel_ms["convert_direct<rgbaf32,rgba8>"].append(events[1].elapsed_time(events[2]))

    def roofline(self, peak, peak_kind):
        # dominant kernel = the slower direction; algorithmic bytes per launch = 20 B/px * 8192^2
        avg = {k: float(np.mean(v)) for k, v in self.kernel_ms.items() if v}
        k = max(avg, key=avg.get)
        alg = self.bytes_per_px * self.W * self.H
        ach = alg / (avg[k] * 1e-3) / 1e9
        other = {kk: round(alg / (vv * 1e-3) / 1e9, 1) for kk, vv in avg.items()}
        return {"bound": "hbm", "kernel": k, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                "frac": round(ach / peak, 4), "peak_kind": peak_kind, "traffic": None,
                "algorithmic_bytes_per_launch": alg, "avg_launch_ms": round(avg[k], 4),
                "all_kernels_GBps": other}

    # -- end to end through the host-pointer C ABI (pinned host buffers, copies inside the timed region)
    def e2e_setup(self):
        import ctypes as C
        W, H = self.W, self.H
        L = self.L
        self.h_u8 = L.gb200_host_alloc(W * H * 4)
        self.h_f32 = L.gb200_host_alloc(W * H * 16)
        if not self.h_u8 or not self.h_f32:
            raise RuntimeError("pinned alloc failed")
        a = np.ctypeslib.as_array(C.cast(self.h_u8, C.POINTER(C.c_uint8)), shape=(W * H * 4,))
        a[:] = np.random.default_rng(1).integers(0, 256, W * H * 4, dtype=np.uint8)
        self.h2d = W * H * 4 + W * H * 16
        self.d2h = W * H * 16 + W * H * 4

    def e2e_step(self):
        from gamut_b200.types import PixelType as PT
        W, H, L = self.W, self.H, self.L
        ok1 = L.gb200_scanlines_convert(PT.rgba8, self.h_u8, W * 4, PT.rgbaf32, self.h_f32, W * 16, W, H)
        ok2 = L.gb200_scanlines_convert(PT.rgbaf32, self.h_f32, W * 16, PT.rgba8, self.h_u8, W * 4, W, H)
        if not (ok1 and ok2):
            raise RuntimeError(self.L.gb200_last_error().decode())

    # -- CPU oracle on a bounded sample of the same workload
    @staticmethod
    def cpu_run(threads: int, rows: int, reps: int):
        from oracle import pyoracle
        from gamut_b200.types import PixelType as PT
        W = ConvertWorkload.W
        rng = np.random.default_rng(1)
        u8 = rng.integers(0, 256, rows * W * 4, dtype=np.uint8)
        f = np.zeros(rows * W * 16, np.uint8)
        back = np.zeros(rows * W * 4, np.uint8)
        pyoracle.lib()
        per = (rows + threads - 1) // threads

        def work(t):
            r0 = t * per
            r1 = min(rows, r0 + per)
            if r1 <= r0:
                return
            n = r1 - r0
            pyoracle.scanlines_convert(PT.rgba8, u8, W * 4, PT.rgbaf32, f, W * 16, W, n, src_off=r0 * W * 4, dst_off=r0 * W * 16)
            pyoracle.scanlines_convert(PT.rgbaf32, f, W * 16, PT.rgba8, back, W * 4, W, n, src_off=r0 * W * 16, dst_off=r0 * W * 4)

        def one():
            if threads == 1:
                work(0)
            else:
                ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
                [t.start() for t in ts]
                [t.join() for t in ts]

        one()  # warm-up
        times = []
        for _ in range(reps):
            t0 = time.perf_counter(); one(); times.append(time.perf_counter() - t0)
        px = 2 * rows * W
        assert np.array_equal(back, u8)
        return px, times


WORKLOADS = {"convert": ConvertWorkload}


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (the C oracle port -- the D
    reference cannot be compiled in this image) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    rows = 2048
    px, times = wl.cpu_run(cores, rows, args.warmup + args.steps)
    times = times[args.warmup:]
    t = float(np.mean(times))
    v = px / t / 1e6
    out = {"metric": "Mpixels/s", "value": round(v, 1), "unit": "Mpixels/s", "impl": "reference",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t * 1e3, 3),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
           "config": {"workload": wl.name},
           "cpu_baseline": {"value": round(v, 1), "unit": "Mpixels/s", "cores": cores, "kind": "port",
                            "sample": f"{rows} rows of the 8192-wide image, both directions, {cores} threads "
                                      "(C restatement of scanline.d; no D toolchain in the image)"},
           "e2e": {"value": round(v, 1), "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="convert", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from gamut_b200 import _lib
    L = _lib.lib()
    if not L.gb200_init():
        raise SystemExit("bench.py: " + L.gb200_last_error().decode())

    wl = WORKLOADS[args.workload](rank)
    stream = torch.cuda.current_stream()
    peak, peak_kind = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        wl.step(stream)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n_ev = 3
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(n_ev)] for _ in range(args.steps)]
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    launches0 = L.gb200_launch_count()
    barrier()
    t_start.record(stream)
    for i in range(args.steps):
        wl.step(stream, evs[i])
    t_end.record(stream)
    barrier()
    launches = L.gb200_launch_count() - launches0
    ms = t_start.elapsed_time(t_end)
    for e in evs:
        wl.collect(e)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the host-pointer C ABI
    wl.e2e_setup()
    wl.e2e_step()  # warm-up (allocates the cached device buffers)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        wl.e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    if rank == 0:
        total_px = wl.px_per_step * args.steps * world
        value = total_px / (ms * 1e-3) / 1e6
        e2e_v = wl.px_per_step * args.e2e_steps * world / e2e_s / 1e6
        out = {"metric": "Mpixels/s", "value": round(value, 1), "unit": "Mpixels/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4),
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype,
               "data": "synthetic",
               "config": {"workload": wl.name, "units_per_rank": "1 image 8192x8192, forward+reverse",
                          "l2": "inputs larger than L2 (256 MiB / 1 GiB per launch, separate buffers per direction)",
                          "sharding": "one image per rank, no collective on the data path"},
               "roofline": wl.roofline(peak, peak_kind),
               "e2e": {"value": round(e2e_v, 1), "unit": "Mpixels/s", "h2d_bytes_per_step": wl.h2d,
                       "d2h_bytes_per_step": wl.d2h, "steps": args.e2e_steps,
                       "api": "gb200_scanlines_convert (host pointers, pinned)"},
               "gpu_launches": int(launches), "clocks": clocks}
        if not args.no_cpu_baseline and world >= 1:
            px, times = wl.cpu_run(1, 512, 3)
            out["cpu_baseline"] = {"value": round(px / float(np.mean(times)) / 1e6, 1), "unit": "Mpixels/s",
                                   "cores": 1, "kind": "port",
                                   "sample": "512 rows of the 8192-wide image, both directions, 1 thread, mean of 3 "
                                             "(C restatement of scanline.d; the reference library is single-threaded)"}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
