#!/usr/bin/env python
"""bench.py -- hot-path throughput on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload jpeg|png|qoix|convert|qoi|bmp|qoix_encode]
                    [--impl reference] [--only]

A "step" is one pass of the hot path over one batch of synthetic input. The DEFAULT workload is the decode
headline of BASELINE.json, configs[3]: JPEG baseline decode (Huffman + IDCT + YCbCr) of a TOTAL batch of 4096
3840x2160 4:2:0 images, strong-scaled: the batch is cut into contiguous index ranges over the ranks
(gamut_b200/shard.py), one process per GPU (torchrun for N>1), no collective on the data path. Every rank
walks its share in sub-batches so that outputs (102 GB for the whole batch) fit HBM.

With no --workload/--only the line of the primary workload also carries, under detail.workloads, full lines
(value, roofline, cpu_baseline, e2e) of the other configs that run on a GPU: configs[1] convert 8192x8192,
configs[2] PNG 1024x1080p, configs[4] QOIX 10-bit + LZ4 (256 images per GPU) and configs[0] QOI 512x512.

Prints ONE JSON line on rank 0. Workload classes live in benchlib.py.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_BEGIN = time.perf_counter()
SECONDARY_BUDGET_S = float(os.environ.get("GB200_BENCH_BUDGET_S", "420"))   # stop adding secondary workloads after this


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
            names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

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
    """--impl reference: the reference's CPU implementation of the path on all host threads. The D reference
    cannot be compiled in this image (no D toolchain), so this is the C restatement under oracle/ (kind "port").
    Each step decodes/converts a bounded sample of the workload (one unit per host thread)."""
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
           "higher_is_better": True, "scaling": getattr(wl, "scaling", "weak"), "vs_baseline": None, "dtype": wl.dtype,
           "data": "synthetic", "config": {"workload": wl.name},
           "cpu_baseline": {"value": round(v, 1), "unit": "Mpixels/s", "cores": cores, "kind": "port",
                            "sample": sample + " (C restatement of the reference; no D toolchain in the image)"},
           "e2e": {"value": round(v, 1), "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


class Ctx:
    """Process-wide state shared by the workloads of one bench run."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        from gamut_b200 import _lib
        self.L = _lib.lib()
        if not self.L.gb200_init():
            raise SystemExit("bench.py: " + self.L.gb200_last_error().decode())
        self.peak, self.peak_kind = load_peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def allsum(self, vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]


def pcie_probe(ctx: Ctx):
    """What the platform gives a plain pinned copy, all ranks at once: the ceiling of any end-to-end number.
    {"d2h": aggregate GB/s, "h2d": aggregate GB/s} over 1 GiB per rank and direction (best of 3)."""
    torch = ctx.torch
    n = 1 << 30
    try:
        h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        d = torch.empty(n, dtype=torch.uint8, device="cuda")
        out = {}
        for name, (dst, src) in (("d2h", (h, d)), ("h2d", (d, h))):
            best = 1e9
            for _ in range(3):
                ctx.barrier()
                a = torch.cuda.Event(enable_timing=True)
                b = torch.cuda.Event(enable_timing=True)
                a.record()
                dst.copy_(src, non_blocking=True)
                b.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            best = ctx.allmax([best])[0]
            out[name] = round(n * ctx.world / (best * 1e-3) / 1e9, 1)
        del h, d
        return out
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:100]}


def run_workload(WL, ctx: Ctx, args, steps, warmup, e2e_steps, cpu_baseline=True):
    """Measures one workload on all ranks; returns the JSON line (rank 0) or None."""
    torch, L = ctx.torch, ctx.L
    wl = WL(ctx.rank, ctx.world, args)
    stream = torch.cuda.current_stream()
    for _ in range(warmup):
        wl.step(stream, timed=False)
    ctx.barrier()
    sampler = ClockSampler(ctx.local)
    if ctx.rank == 0:
        sampler.start()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    launches0 = L.gb200_launch_count()
    ctx.barrier()
    t_start.record(stream)
    for _ in range(steps):
        wl.step(stream, timed=True)
    t_end.record(stream)
    ctx.barrier()
    launches = L.gb200_launch_count() - launches0
    ms = t_start.elapsed_time(t_end)
    wl.finish_timing()
    ms = ctx.allmax([ms])[0]
    clocks = sampler.stop() if ctx.rank == 0 else None

    # ---- end to end through the host-pointer C ABI: every step starts with its inputs in (pinned) host memory
    # and ends with its results in host memory (H2D and D2H inside the timed region)
    e2e = None
    pcie = None
    if e2e_steps > 0:
        pcie = pcie_probe(ctx)
        wl.e2e_setup()
        wl.e2e_step()                      # warm-up (allocates the cached device / pinned buffers)
        wl.e2e_step()
        ctx.barrier()
        t0 = time.perf_counter()
        step_ms = []
        for _ in range(e2e_steps):
            t1 = time.perf_counter()
            wl.e2e_step()
            step_ms.append(round((time.perf_counter() - t1) * 1e3, 2))
        torch.cuda.synchronize()
        total_s = time.perf_counter() - t0
        median_s = float(np.median(step_ms)) * 1e-3 * e2e_steps
        median_s, total_s = ctx.allmax([median_s, total_s])
        e2e = (median_s, total_s, step_ms)
        wl.e2e_teardown()

    px_all, e2e_px_all = ctx.allsum([wl.px_per_step, wl.e2e_px_per_step])
    line = None
    if ctx.rank == 0:
        value = px_all * steps / (ms * 1e-3) / 1e6
        cfg = {"workload": wl.name, "sharding": "units sharded across ranks, no collective on the data path"}
        cfg.update(wl.config())
        line = {"metric": "Mpixels/s", "value": round(value, 1), "unit": "Mpixels/s", "n_gpus": ctx.world,
                "steps": steps, "warmup": warmup, "ms_per_step": round(ms / steps, 4),
                "higher_is_better": True, "scaling": getattr(wl, "scaling", "weak"), "vs_baseline": None,
                "dtype": wl.dtype, "data": "synthetic", "config": cfg,
                "roofline": wl.roofline(ctx.peak, ctx.peak_kind)}
        if e2e:
            median_s, total_s, step_ms = e2e
            line["e2e"] = {"value": round(e2e_px_all * e2e_steps / total_s / 1e6, 1), "unit": "Mpixels/s",
                           "h2d_bytes_per_step": wl.h2d, "d2h_bytes_per_step": wl.d2h, "steps": e2e_steps,
                           "timing": "mean over the steps (wall clock around the whole loop, max over ranks)",
                           "value_median": round(e2e_px_all * e2e_steps / median_s / 1e6, 1),
                           "step_ms": step_ms, "api": wl.e2e_api, "platform_copy_GBps": pcie}
        line["gpu_launches"] = int(launches)
        line["clocks"] = clocks
        extra = wl.extra()
        if extra:
            line["detail"] = extra
        if cpu_baseline:
            px, times, sample = wl.cpu_run(1, 6, full=False)
            times = times[1:]
            line["cpu_baseline"] = {"value": round(px / float(np.median(times)) / 1e6, 1), "unit": "Mpixels/s",
                                    "cores": 1, "kind": "port", "timing": "1 warm-up, median of 5",
                                    "sample": sample + " (C restatement of the reference, which is single-threaded)"}
    wl.release()
    del wl
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    L.gb200_device_trim()
    return line


def main():
    from benchlib import WORKLOADS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None, help="TOTAL images of the batched decode workloads")
    ap.add_argument("--sub-batch", type=int, default=None, help="images per decode call inside a step")
    ap.add_argument("--only", action="store_true", help="primary workload only (no detail.workloads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    secondary = args.workload is None and not args.only
    if args.workload is None:
        args.workload = "jpeg"
    WL = WORKLOADS[args.workload]
    if args.steps is None:
        args.steps = WL.default_steps if args.impl == "b200" else 3
    if args.impl == "reference":
        reference_arm(args, WORKLOADS)
        return
    if args.warmup < 3:
        sys.stderr.write("bench.py: --warmup %d raised to 3 (timing rules: >= 3 warm-up steps)\n" % args.warmup)
        args.warmup = 3

    ctx = Ctx(args)
    line = run_workload(WL, ctx, args, args.steps, args.warmup,
                        WL.default_e2e_steps if args.e2e_steps is None else args.e2e_steps,
                        cpu_baseline=not args.no_cpu_baseline)
    if secondary:
        others = {}
        # at N > 1 only the workload whose BASELINE config is multi-GPU (QOIX, configs[4]) rides along
        for name in (("convert", "png", "qoix", "qoi", "bmp", "qoix_encode") if ctx.world == 1 else ("qoix",)):
            if ctx.allmax([time.perf_counter() - T_BEGIN])[0] > SECONDARY_BUDGET_S:
                others[name] = {"skipped": "time budget of the default run (%.0f s) used up" % SECONDARY_BUDGET_S}
                continue
            W2 = WORKLOADS[name]
            a2 = argparse.Namespace(**vars(args))
            a2.batch = None
            a2.sub_batch = None
            try:
                l2 = run_workload(W2, ctx, a2, W2.default_steps, 3, W2.default_e2e_steps, cpu_baseline=not args.no_cpu_baseline)
            except Exception as e:  # noqa: BLE001 -- a secondary workload never costs the primary line
                l2 = {"error": "%s: %s" % (type(e).__name__, e)}
            others[name] = l2
        if line is not None:
            line.setdefault("detail", {})["workloads"] = others
    if ctx.rank == 0:
        line["bench_wall_s"] = round(time.perf_counter() - T_BEGIN, 1)
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
