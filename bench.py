#!/usr/bin/env python
"""bench.py -- hot-path throughput on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload convert|png|jpeg|qoix] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input. Default workload is
BASELINE.json configs[1]: PixelType convert rgba8<->rgbaf32 on one 8192x8192 image (forward +
reverse = 2 launches, 134.2 Mpixels per step). One process per GPU (torchrun for N>1), batch units
sharded across ranks with no data-path collective ("weak" scaling: every rank converts its own image).
Workload classes live in benchlib.py.

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



def reference_arm(args, WORKLOADS):
    """--impl reference: the reference's CPU implementation of the path (the C oracle port -- the D
    reference cannot be compiled in this image) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    px, times, sample = wl.cpu_run(cores, args.warmup + args.steps, full=True)
    times = times[args.warmup:]
    t = float(np.mean(times))
    v = px / t / 1e6
    out = {"metric": "Mpixels/s", "value": round(v, 1), "unit": "Mpixels/s", "impl": "reference",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t * 1e3, 3),
           "higher_is_better": True, "scaling": getattr(wl, "scaling", "weak"), "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
           "config": {"workload": wl.name},
           "cpu_baseline": {"value": round(v, 1), "unit": "Mpixels/s", "cores": cores, "kind": "port",
                            "sample": sample + " (C restatement of the reference; no D toolchain in the image)"},
           "e2e": {"value": round(v, 1), "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    from benchlib import WORKLOADS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="convert", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None, help="images per rank for the batched decode workloads")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    WL = WORKLOADS[args.workload]
    if args.steps is None:
        args.steps = WL.default_steps if args.impl == "b200" else 3
    if args.e2e_steps is None:
        args.e2e_steps = WL.default_e2e_steps

    if args.impl == "reference":
        reference_arm(args, WORKLOADS)
        return
    args.warmup = max(args.warmup, 3)

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

    wl = WL(rank, world, args)
    stream = torch.cuda.current_stream()
    peak, peak_kind = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        wl.step(stream, timed=False)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    launches0 = L.gb200_launch_count()
    barrier()
    t_start.record(stream)
    for i in range(args.steps):
        wl.step(stream, timed=True)
    t_end.record(stream)
    barrier()
    launches = L.gb200_launch_count() - launches0
    ms = t_start.elapsed_time(t_end)
    wl.finish_timing()
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
    e2e_step_ms = []
    for _ in range(args.e2e_steps):
        t1 = time.perf_counter()
        wl.e2e_step()
        e2e_step_ms.append(round((time.perf_counter() - t1) * 1e3, 2))
    torch.cuda.synchronize()
    e2e_total_s = time.perf_counter() - t0
    # every e2e step ends with its results in host memory (the C ABI calls synchronise), so steps are timed one by
    # one and the median step is reported (median of >= 5): a fresh box shows occasional 2-3x spikes from
    # host-side noise, visible in "step_ms"; "value_mean" keeps the plain total/steps figure
    e2e_s = float(np.median(e2e_step_ms)) * 1e-3 * args.e2e_steps if e2e_step_ms else e2e_total_s
    if world > 1:
        t = torch.tensor([e2e_s, e2e_total_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, e2e_total_s = float(t[0].item()), float(t[1].item())

    # units processed by all ranks (ranges are balanced by bytes, so the per-rank counts may differ by one image)
    px_all, e2e_px_all = wl.px_per_step * world, wl.e2e_px_per_step * world
    if world > 1:
        t = torch.tensor([wl.px_per_step, wl.e2e_px_per_step], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        px_all, e2e_px_all = float(t[0].item()), float(t[1].item())
    if rank == 0:
        total_px = px_all * args.steps
        value = total_px / (ms * 1e-3) / 1e6
        e2e_v = e2e_px_all * args.e2e_steps / e2e_s / 1e6
        cfg = {"workload": wl.name, "sharding": "units sharded across ranks, no collective on the data path"}
        cfg.update(wl.config())
        out = {"metric": "Mpixels/s", "value": round(value, 1), "unit": "Mpixels/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4),
               "higher_is_better": True, "scaling": getattr(wl, "scaling", "weak"), "vs_baseline": None, "dtype": wl.dtype,
               "data": "synthetic", "config": cfg,
               "roofline": wl.roofline(peak, peak_kind),
               "e2e": {"value": round(e2e_v, 1), "unit": "Mpixels/s", "h2d_bytes_per_step": wl.h2d,
                       "d2h_bytes_per_step": wl.d2h, "steps": args.e2e_steps, "timing": "median step",
                       "value_mean": round(e2e_px_all * args.e2e_steps / e2e_total_s / 1e6, 1), "step_ms": e2e_step_ms, "api": wl.e2e_api},
               "gpu_launches": int(launches), "clocks": clocks}
        extra = wl.extra()
        if extra:
            out["detail"] = extra
        if not args.no_cpu_baseline:
            px, times, sample = wl.cpu_run(1, 3, full=False)
            out["cpu_baseline"] = {"value": round(px / float(np.mean(times)) / 1e6, 1), "unit": "Mpixels/s",
                                   "cores": 1, "kind": "port",
                                   "sample": sample + " (C restatement of the reference, which is single-threaded)"}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
