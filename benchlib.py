"""Workloads for bench.py (one class per BASELINE.json config that runs on the GPU).

Every workload keeps its inputs resident in HBM for `step()` (the `value` leg), offers an end-to-end
leg through the host-pointer C ABI (`e2e_*`), and a CPU leg that times the oracle port on a bounded
sample of the same workload (`cpu_run`; the only place besides tests/ and smoke() that uses oracle/)."""
from __future__ import annotations

import ctypes as C
import io
import os
import threading
import time

import numpy as np


ROOT = os.path.dirname(os.path.abspath(__file__))


def load_traffic():
    """Per-kernel DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of THIS build, captured by
    scripts/capture_traffic.py (one ncu pass per workload) into profiles/traffic.json; None when not captured."""
    import json
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


def hbm_peak():
    import json
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0      # fallback of B200_PROFILING.md


def traffic_for(kernel_key, units):
    """(bytes per launch at `units` units per launch, source string) for a kernel of profiles/traffic.json."""
    t = load_traffic().get(kernel_key)
    if not t:
        return None, "not captured for this build (scripts/capture_traffic.py)"
    return int(t["bytes_per_unit"] * units), t.get("source", "profiles/traffic.json")


class _WorkloadBase:
    scaling = "weak"

    def finish_timing(self):
        pass

    def extra(self):
        return None

    def e2e_teardown(self):
        """Returns the pinned host buffers of the e2e leg."""
        for name in list(getattr(self, "_pinned", [])):
            self.L_free(name)
        self._pinned = []

    def pin(self, nbytes):
        from gamut_b200 import _lib
        p = _lib.lib().gb200_host_alloc(nbytes)
        if not p:
            raise RuntimeError("pinned alloc of %d bytes failed" % nbytes)
        self.__dict__.setdefault("_pinned", []).append(p)
        return p

    def L_free(self, p):
        from gamut_b200 import _lib
        _lib.lib().gb200_host_free(p)

    def release(self):
        """Drops device tensors so that the next workload of the same process starts with an empty HBM."""
        for k, v in list(self.__dict__.items()):
            if k not in ("px_per_step", "e2e_px_per_step", "h2d", "d2h"):
                self.__dict__.pop(k, None)


def _threads_run(fn, threads):
    if threads == 1:
        fn(0)
        return
    ts = [threading.Thread(target=fn, args=(t,)) for t in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]


def _e2e_threads(n, fn):
    """fn(a, b) over GB200_E2E_THREADS (default 1) contiguous slices of range(n), one host thread each."""
    parts = max(1, min(int(os.environ.get("GB200_E2E_THREADS", "1")), n))
    if parts == 1:
        fn(0, n)
        return
    cuts = [n * i // parts for i in range(parts + 1)]
    err = []

    def run(t):
        try:
            fn(cuts[t], cuts[t + 1])
        except BaseException as e:  # noqa: BLE001
            err.append(e)

    _threads_run(run, parts)
    if err:
        raise err[0]


class EventLog:
    """Collects (tag, start_event, end_event) on torch's current stream; elapsed read after the sync."""

    def __init__(self):
        import torch
        self.torch = torch
        self.items = []

    def span(self, tag, stream):
        a = self.torch.cuda.Event(enable_timing=True)
        b = self.torch.cuda.Event(enable_timing=True)
        self.items.append((tag, a, b))
        a.record(stream)
        return b

    def collect(self):
        out = {}
        for tag, a, b in self.items:
            out.setdefault(tag, []).append(a.elapsed_time(b))
        return out


# ----------------------------------------------------------------------------------------------
class ConvertWorkload(_WorkloadBase):
    """BASELINE configs[1]: PixelType convert rgba8 <-> rgbaf32, 8192x8192, 1 GPU (HBM roofline probe)."""
    name = "PixelType convert rgba8<->rgbaf32 8192x8192 (BASELINE configs[1])"
    dtype = "f32"
    default_steps = 100
    default_e2e_steps = 9
    W = H = 8192
    bytes_per_px = 20            # 4 B rgba8 + 16 B rgbaf32 per direction
    e2e_api = ("gb200_scanlines_convert (host pointers, pinned); forward and reverse issued concurrently from two "
               "host threads (the entry point is thread-safe), each call pipelines H2D/kernel/D2H over row bands")

    def __init__(self, rank, world, args):
        import torch
        from gamut_b200 import _lib
        self.L = _lib.lib()
        W, H = self.W, self.H
        g = torch.Generator(device="cuda").manual_seed(1 + rank)
        self.u8 = torch.randint(0, 256, (H, W, 4), dtype=torch.uint8, device="cuda", generator=g)
        self.f32 = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
        self.f32_in = torch.rand((H, W, 4), dtype=torch.float32, device="cuda", generator=g)
        self.u8_out = torch.empty_like(self.u8)
        self.px_per_step = 2 * W * H
        self.e2e_px_per_step = 2 * W * H
        self.log = EventLog()
        self.kernel_ms = {}

    def step(self, stream, timed):
        from gamut_b200.types import PixelType as PT
        W, H, L = self.W, self.H, self.L
        st = stream.cuda_stream
        e = self.log.span("convert_direct<rgba8,rgbaf32>", stream) if timed else None
        ok1 = L.gb200_scanlines_convert_device(PT.rgba8, self.u8.data_ptr(), W * 4, PT.rgbaf32, self.f32.data_ptr(), W * 16, W, H, st)
        if e:
            e.record(stream)
        e = self.log.span("convert_direct<rgbaf32,rgba8>", stream) if timed else None
        ok2 = L.gb200_scanlines_convert_device(PT.rgbaf32, self.f32_in.data_ptr(), W * 16, PT.rgba8, self.u8_out.data_ptr(), W * 4, W, H, st)
        if e:
            e.record(stream)
        if not (ok1 and ok2):
            raise RuntimeError(L.gb200_last_error().decode())

    def finish_timing(self):
        self.kernel_ms = self.log.collect()

    def config(self):
        return {"units_per_rank": "1 image 8192x8192, forward+reverse",
                "l2": "inputs larger than L2 (256 MiB / 1 GiB per launch, separate buffers per direction)"}

    def roofline(self, peak, peak_kind):
        avg = {k: float(np.mean(v)) for k, v in self.kernel_ms.items() if v}
        k = max(avg, key=avg.get)
        alg = self.bytes_per_px * self.W * self.H
        ach = alg / (avg[k] * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": k, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                "frac": round(ach / peak, 4), "peak_kind": peak_kind,
                "traffic": traffic_for(k, 1)[0], "traffic_source": traffic_for(k, 1)[1],
                "algorithmic_bytes_per_launch": alg, "avg_launch_ms": round(avg[k], 4),
                "all_kernels_GBps": {kk: round(alg / (vv * 1e-3) / 1e9, 1) for kk, vv in avg.items()}}

    def e2e_setup(self):
        W, H, L = self.W, self.H, self.L
        self.h_u8 = self.pin(W * H * 4)
        self.h_f32 = self.pin(W * H * 16)
        a = np.ctypeslib.as_array(C.cast(self.h_u8, C.POINTER(C.c_uint8)), shape=(W * H * 4,))
        a[:] = np.random.default_rng(1).integers(0, 256, W * H * 4, dtype=np.uint8)
        # second pair of buffers for the reverse direction so that both directions can run concurrently
        self.h_f32_in = self.pin(W * H * 16)
        self.h_u8_out = self.pin(W * H * 4)
        assert L.gb200_scanlines_convert(PT_rgba8(), self.h_u8, W * 4, PT_rgbaf32(), self.h_f32_in, W * 16, W, H)
        self.h2d = W * H * 4 + W * H * 16
        self.d2h = W * H * 16 + W * H * 4

    def e2e_step(self):
        from gamut_b200.types import PixelType as PT
        W, H, L = self.W, self.H, self.L
        ok = [0, 0]

        def run(t):
            if t == 0:
                ok[0] = L.gb200_scanlines_convert(PT.rgba8, self.h_u8, W * 4, PT.rgbaf32, self.h_f32, W * 16, W, H)
            else:
                ok[1] = L.gb200_scanlines_convert(PT.rgbaf32, self.h_f32_in, W * 16, PT.rgba8, self.h_u8_out, W * 4, W, H)

        _threads_run(run, 2)
        if not (ok[0] and ok[1]):
            raise RuntimeError("gb200_scanlines_convert failed")
        if not getattr(self, "_e2e_checked", False):
            a = np.ctypeslib.as_array(C.cast(self.h_u8, C.POINTER(C.c_uint8)), shape=(W * H * 4,))
            b = np.ctypeslib.as_array(C.cast(self.h_u8_out, C.POINTER(C.c_uint8)), shape=(W * H * 4,))
            assert np.array_equal(a, b), "e2e round trip rgba8 -> rgbaf32 -> rgba8 is not the identity"
            self._e2e_checked = True

    @staticmethod
    def cpu_run(threads, reps, full):
        from oracle import pyoracle
        from gamut_b200.types import PixelType as PT
        W = ConvertWorkload.W
        rows = 2048 if full else 512
        rng = np.random.default_rng(1)
        u8 = rng.integers(0, 256, rows * W * 4, dtype=np.uint8)
        f = np.zeros(rows * W * 16, np.uint8)
        back = np.zeros(rows * W * 4, np.uint8)
        pyoracle.lib()
        per = (rows + threads - 1) // threads

        def work(t):
            r0 = t * per
            n = min(rows, r0 + per) - r0
            if n <= 0:
                return
            pyoracle.scanlines_convert(PT.rgba8, u8, W * 4, PT.rgbaf32, f, W * 16, W, n, src_off=r0 * W * 4, dst_off=r0 * W * 16)
            pyoracle.scanlines_convert(PT.rgbaf32, f, W * 16, PT.rgba8, back, W * 4, W, n, src_off=r0 * W * 16, dst_off=r0 * W * 4)

        _threads_run(work, threads)
        times = []
        for _ in range(reps):
            t0 = time.perf_counter()
            _threads_run(work, threads)
            times.append(time.perf_counter() - t0)
        assert np.array_equal(back, u8)
        return 2 * rows * W, times, f"{rows} rows of the 8192-wide image, both directions, {threads} thread(s)"


# ----------------------------------------------------------------------------------------------
def synth_photo(h, w, c, seed):
    """Photo-like synthetic image: low-frequency sinusoids + noise (+ alpha gradient). uint8 (h, w, c)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, c), np.float32)
    for k in range(min(c, 3)):
        a = np.zeros((h, w), np.float32)
        for _ in range(4):
            fx, fy, ph = rng.uniform(0.002, 0.03), rng.uniform(0.002, 0.03), rng.uniform(0, 6.28)
            a += np.sin(xx * fx + yy * fy + ph) * rng.uniform(0.1, 0.3)
        img[:, :, k] = 0.5 + a * 0.5
    img[:, :, :min(c, 3)] += rng.normal(0, 0.012, (h, w, min(c, 3))).astype(np.float32)
    if c in (2, 4):
        img[:, :, c - 1] = np.clip((xx + yy) / (h + w) * 1.3, 0, 1)
    return (np.clip(img, 0, 1) * 255 + 0.5).astype(np.uint8)


def synth_photo_mixed(h, w, c, seed):
    """Photo-like synthetic image with varied content: a smooth base plus bands of different structure (noisy smooth
    areas, vertical and horizontal structures, clean diagonal edges, clean blobs, fine texture). Which PNG filter
    wins a row depends on the structure, so an adaptive encoder emits Sub, Up, Avg and Paeth rows like it does on
    photographs (synth_photo alone makes every row pick the same filter)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    nc = min(c, 3)
    img = np.zeros((h, w, c), np.float32)
    for k in range(nc):
        a = np.zeros((h, w), np.float32)
        for _ in range(4):
            fx, fy, ph = rng.uniform(0.002, 0.03), rng.uniform(0.002, 0.03), rng.uniform(0, 6.28)
            a += np.sin(xx * fx + yy * fy + ph) * rng.uniform(0.1, 0.3)
        img[:, :, k] = 0.5 + a * 0.4
    y = 0
    while y < h:
        bh = int(rng.integers(24, 120))
        kind = int(rng.integers(0, 6))
        sl = slice(y, min(h, y + bh))
        X, Y = xx[sl], yy[sl]
        amp = rng.uniform(0.08, 0.25)
        sigma = [0.008, 0.004, 0.004, 0.0, 0.0, 0.0][kind]
        for k in range(nc):
            if kind == 0:
                t = 0
            elif kind == 1:      # vertical structures: constant down the rows
                col = np.cumsum(rng.normal(0, 0.6, w)).astype(np.float32)
                t = np.sign(np.sin(col))[None, :] * amp
            elif kind == 2:      # horizontal structures: constant along the row
                row = np.cumsum(rng.normal(0, 0.6, X.shape[0])).astype(np.float32)
                t = np.sign(np.sin(row))[:, None] * amp
            elif kind == 3:      # clean diagonal hard edges
                f = rng.uniform(0.05, 0.2)
                t = np.sign(np.sin((X + Y * rng.choice([-1.0, 1.0])) * f + k)) * amp
            elif kind == 4:      # clean blobs: curved edges in every direction
                t = np.zeros_like(X)
                for _ in range(12):
                    cx, cy, r = rng.uniform(0, w), rng.uniform(y, y + bh), rng.uniform(10, 60)
                    t += (np.hypot(X - cx, Y - cy) < r) * rng.uniform(-amp, amp)
            else:                # fine isotropic texture
                t = rng.normal(0, amp * 0.6, X.shape).astype(np.float32)
            img[sl, :, k] += t
        if sigma:
            img[sl, :, :nc] += rng.normal(0, sigma, (X.shape[0], w, nc)).astype(np.float32)
        y += bh
    if c in (2, 4):
        img[:, :, c - 1] = np.clip((xx + yy) / (h + w) * 1.3, 0, 1)
    return (np.clip(img, 0, 1) * 255 + 0.5).astype(np.uint8)


def make_png_files(distinct, w, h, seed0=1000):
    """RGBA8 PNG files of mixed-content images, filters chosen per row by libpng's minimum-sum-of-absolute-
    differences heuristic (tests/pngwriter.py), zlib level 6."""
    sys_path_tests()
    from pngwriter import write_png
    return [write_png(synth_photo_mixed(h, w, 4, seed0 + i), 6, 8, filters="adaptive", level=6) for i in range(distinct)]


def png_filter_histogram(files):
    """Rows per filter type (None, Sub, Up, Avg, Paeth) over the given PNG files."""
    import struct
    import zlib
    hist = np.zeros(5, np.int64)
    for f in files:
        w, h = struct.unpack(">II", f[16:24])
        raw = np.frombuffer(zlib.decompress(split_idat(f)), np.uint8).reshape(h, -1)
        hist += np.bincount(raw[:, 0], minlength=5)[:5]
    return [int(x) for x in hist]


def split_idat(png: bytes):
    """Concatenated IDAT payload of a PNG file (host-side helper for the kernel-only legs)."""
    import struct
    pos, out = 8, b""
    while pos + 8 <= len(png):
        n, typ = struct.unpack(">I4s", png[pos:pos + 8])
        if typ == b"IDAT":
            out += png[pos + 8:pos + 8 + n]
        pos += 12 + n
    return out


class PngWorkload(_WorkloadBase):
    """BASELINE configs[2]: PNG 8-bit RGBA decode (inflate + unfilter), batch of 1920x1080 images, 1 GPU."""
    name = "PNG 8-bit RGBA decode + unfilter, batch 1024 images 1920x1080 (BASELINE configs[2])"
    dtype = "u8"
    default_steps = 3
    default_e2e_steps = 5
    W, H = 1920, 1080
    DISTINCT = 8
    e2e_api = "gb200_decode_batch_host(PNG): host file bytes in, rgba8 pixels in pinned host memory out; sub-batches pipelined (download of k overlaps upload + kernels of k+1)"

    def __init__(self, rank, world, args):
        import torch
        from gamut_b200 import codecs
        self.torch = torch
        self.codecs = codecs
        self.n = args.batch or 1024
        self.files = make_png_files(self.DISTINCT, self.W, self.H, 1000 + 100 * rank)
        self.host_files = [self.files[i % self.DISTINCT] for i in range(self.n)]
        # distinct device copies of every file so that nothing is served from L2 by aliasing
        base = [torch.frombuffer(bytearray(f + b"\0" * 64), dtype=torch.uint8).cuda() for f in self.files]
        self.dev_bufs = [base[i % self.DISTINCT].clone() for i in range(self.n)]
        self.dev_ptrs = [t.data_ptr() for t in self.dev_bufs]
        self.px_per_step = self.n * self.W * self.H
        self.e2e_n = min(self.n, 512)
        self.e2e_px_per_step = self.e2e_n * self.W * self.H
        self.comp_bytes = sum(len(split_idat(f)) for f in self.files) / self.DISTINCT
        self.phase = []
        self.unf_ms = []
        self.log = EventLog()
        # raw (inflated) streams resident in HBM for the unfilter-only leg
        import zlib
        raw = [np.frombuffer(zlib.decompress(split_idat(f)), np.uint8) for f in self.files]
        self.raw_len = raw[0].size
        self.raw_stride = (self.raw_len + 255) // 256 * 256
        rawbuf = np.zeros((self.DISTINCT, self.raw_stride), np.uint8)
        for i in range(self.DISTINCT):
            rawbuf[i, :self.raw_len] = raw[i]
        d16 = torch.from_numpy(rawbuf).cuda()
        self.d_raw = d16[torch.arange(self.n, device="cuda") % self.DISTINCT].contiguous()
        self.out_stride = self.W * self.H * 4
        self.d_unf = torch.empty((self.n, self.out_stride), dtype=torch.uint8, device="cuda")
        # forced-filter variants of the same pixels: every unfilter branch is measured on its own
        sys_path_tests()
        from pngwriter import filter_rows
        self.variants = {}
        nd = 4
        pix = [np.frombuffer(self._unfilter_host(raw[i]), np.uint8).reshape(self.H, self.W * 4) for i in range(nd)]
        self.ref_pixels = torch.from_numpy(pix[0].copy()).cuda()
        for name, filt in (("paeth", 4), ("avg", 3), ("mixed_01234", (0, 1, 2, 3, 4))):
            buf = np.zeros((nd, self.raw_stride), np.uint8)
            for i in range(nd):
                buf[i, :self.raw_len] = np.frombuffer(filter_rows(pix[i], 4, filt), np.uint8)
            d = torch.from_numpy(buf).cuda()
            self.variants[name] = d[torch.arange(self.n, device="cuda") % nd].contiguous()

    def step(self, stream, timed):
        b = self.codecs.png_decode_batch(self.host_files, 0, 0, files_dev=self.dev_ptrs, stream=stream.cuda_stream)
        if timed:
            ph, hp = b.timing()
            self.phase.append(ph[:4] + [hp])
        bad = sum(1 for d in b.images if not d.status)
        b.free()
        if bad:
            raise RuntimeError(f"{bad} PNG images failed to decode")

    def _unfilter_legs(self, stream):
        """Unfilter-only legs on pre-inflated streams (kernel-level roofline). Run AFTER the timed region of the step
        (finish_timing): they are not part of `value`."""
        e = self.log.span("unfilter", stream)
        L = self.codecs._L()
        ok = L.gb200_png_unfilter_device(self.d_raw.data_ptr(), self.raw_stride, self.d_unf.data_ptr(), self.out_stride,
                                         self.n, self.W * 4, self.H, 4, None, stream.cuda_stream)
        e.record(stream)
        assert ok
        for name, d in self.variants.items():
            e = self.log.span("unfilter_" + name, stream)
            ok = L.gb200_png_unfilter_device(d.data_ptr(), self.raw_stride, self.d_unf.data_ptr(), self.out_stride,
                                             self.n, self.W * 4, self.H, 4, None, stream.cuda_stream)
            e.record(stream)
            assert ok
            if not getattr(self, "_checked_" + name, False):
                assert self.torch.equal(self.d_unf[0].view(self.H, self.W * 4), self.ref_pixels), name
                setattr(self, "_checked_" + name, True)

    @staticmethod
    def _unfilter_host(raw):
        """Reference pixels for the variants: undo the PNG filters with the oracle (setup only)."""
        from oracle import pyoracle
        out = pyoracle.png_unfilter(np.ascontiguousarray(raw), 4, 4, PngWorkload.W, PngWorkload.H, 8)
        assert out is not None
        return out.tobytes()

    def finish_timing(self):
        stream = self.torch.cuda.current_stream()
        for _ in range(3):
            self._unfilter_legs(stream)
        stream.synchronize()
        c = self.log.collect()
        self.unf_ms = c.get("unfilter", [])
        self.variant_ms = {k: c.get("unfilter_" + k, []) for k in self.variants}

    def config(self):
        return {"units_per_rank": f"{self.n} images {self.W}x{self.H} RGBA8 ({self.DISTINCT} distinct, zlib level 6, libpng-style adaptive filters)",
                "filter_rows_none_sub_up_avg_paeth": png_filter_histogram(self.files),
                "compressed_idat_bytes_per_image": int(self.comp_bytes),
                "l2": "inputs larger than L2 (every image has its own device copy; batch >> 126 MB)",
                "note": "value = whole decode (gather+inflate+unfilter); the unfilter-only leg is timed separately and excluded"}

    def roofline(self, peak, peak_kind):
        ph = np.mean(np.array(self.phase), axis=0)
        # dominant kernel: inflate. algorithmic bytes per launch = compressed in + raw out, per image * n
        alg = (self.comp_bytes + self.raw_len) * self.n
        ach = alg / (ph[1] * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": "inflate pipeline (infp_find/verify/compact/count/walk/write/resolve kernels)", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                "frac": round(ach / peak, 4), "peak_kind": peak_kind, "traffic": traffic_for("inflate", self.n)[0],
                "traffic_source": traffic_for("inflate", self.n)[1],
                "algorithmic_bytes_per_launch": int(alg), "avg_launch_ms": round(float(ph[1]), 3)}

    def extra(self):
        ph = np.mean(np.array(self.phase), axis=0)
        px = self.n * self.W * self.H
        d = {"phase_ms": {"gather": round(float(ph[0]), 3), "inflate": round(float(ph[1]), 3),
                          "unfilter": round(float(ph[2]), 3), "finish": round(float(ph[3]), 3),
                          "host_parse": round(float(ph[4]), 3)},
             "inflate_only_Mpixels_s": round(px / (ph[1] * 1e-3) / 1e6, 1)}
        if self.unf_ms:
            ms = float(np.mean(self.unf_ms))
            alg = (self.raw_len + self.out_stride) * self.n
            d["unfilter_only"] = {"filters": "the workload's own files (adaptive: Sub/Up/Avg/Paeth rows, see config)",
                                  "Mpixels_s": round(px / (ms * 1e-3) / 1e6, 1), "ms": round(ms, 3),
                                  "algorithmic_bytes": int(alg), "GBps": round(alg / (ms * 1e-3) / 1e9, 1),
                                  "frac_of_measured_hbm": round(alg / (ms * 1e-3) / 1e9 / hbm_peak(), 4)}
            for k, v in getattr(self, "variant_ms", {}).items():
                if v:
                    m2 = float(np.mean(v))
                    d["unfilter_only_" + k] = {"Mpixels_s": round(px / (m2 * 1e-3) / 1e6, 1), "ms": round(m2, 3),
                                               "GBps": round(alg / (m2 * 1e-3) / 1e9, 1),
                                               "frac_of_measured_hbm": round(alg / (m2 * 1e-3) / 1e9 / hbm_peak(), 4)}
        return d

    def e2e_setup(self):
        self.h2d = int(self.comp_bytes * self.e2e_n)
        self.d2h = self.e2e_n * self.out_stride
        self.h_out = self.pin(self.e2e_n * self.out_stride)

    def e2e_step(self):
        # one call for the whole batch: slicing it over several host threads was measured and is slower (the slices
        # serialise on the allocator and each small batch is latency-bound); GB200_E2E_THREADS=k re-measures it
        _e2e_threads(self.e2e_n, self._e2e_slice)

    def _e2e_slice(self, a, b_):
        descs = self.codecs.decode_batch_host(1, self.host_files[a:b_], 0, 0, self.h_out + a * self.out_stride, self.out_stride)
        assert all(d.status for d in descs)

    @staticmethod
    def cpu_run(threads, reps, full):
        from oracle import pyoracle
        W, H = PngWorkload.W, PngWorkload.H
        nimg = max(threads, 2) if full else 2
        files = make_png_files(min(nimg, 4), W, H, 1000)
        pyoracle.lib()

        def work(t):
            for i in range(t, nimg, threads):
                px, _ = pyoracle.png_load(files[i % len(files)], 0, 0)
                assert px is not None

        _threads_run(work, threads)
        times = []
        for _ in range(reps):
            t0 = time.perf_counter()
            _threads_run(work, threads)
            times.append(time.perf_counter() - t0)
        return nimg * W * H, times, f"{nimg} images 1920x1080 RGBA8, {threads} thread(s), one image per worker"



# ----------------------------------------------------------------------------------------------
class _BatchDecodeWorkload(_WorkloadBase):
    """Shared plumbing of the batched decode workloads (JPEG, QOIX): files resident in HBM, one call per step."""
    dtype = "u8"
    scaling = "strong"
    default_steps = 3
    default_e2e_steps = 5
    DISTINCT = 8

    def _setup(self, rank, world, args, total_default):
        import torch
        from gamut_b200 import codecs
        self.torch, self.codecs = torch, codecs
        total = args.batch or total_default
        self.world = world
        # strong scaling: ONE batch of `total` images (image k = distinct file k % DISTINCT, same on every rank) is cut
        # into contiguous index ranges balanced by compressed bytes (gamut_b200/shard.py); no collective
        from gamut_b200 import shard
        self.total = max(total, world)
        self.files = self.make_files(1000)
        nd = len(self.files)
        self.range = shard.my_range([len(self.files[k % nd]) for k in range(self.total)], rank, world)
        idx = list(range(*self.range))
        self.n = len(idx)
        self.host_files = [self.files[k % nd] for k in idx]
        base = [torch.frombuffer(bytearray(f + b"\0" * 64), dtype=torch.uint8).cuda() for f in self.files]
        self.dev_bufs = [base[k % nd].clone() for k in idx]
        self.dev_ptrs = [t.data_ptr() for t in self.dev_bufs]
        self.px_per_step = self.n * self.W * self.H
        self.e2e_n = min(self.n, getattr(self, 'E2E_N', 128))
        self.e2e_px_per_step = self.e2e_n * self.W * self.H
        self.comp_bytes = sum(len(f) for f in self.files) / len(self.files)
        self.sub = min(self.n, args.sub_batch or getattr(self, 'SUB_BATCH', self.n))
        self.phase = []

    def step(self, stream, timed):
        # the rank's share is walked in sub-batches of `sub` images (one decode call each) so that the decoded
        # pixels of a call fit HBM; every call's phase times (CUDA events on the call's stream) are summed
        ph_sum = None
        for a in range(0, self.n, self.sub):
            e = min(self.n, a + self.sub)
            b = self.decode(self.host_files[a:e], self.dev_ptrs[a:e], stream.cuda_stream)
            if timed:
                ph, hp = b.timing()
                row = np.array(ph[:8] + [hp])
                ph_sum = row if ph_sum is None else ph_sum + row
            bad = sum(1 for d in b.images if not d.status)
            b.free()
            if bad:
                raise RuntimeError(f"{bad} images failed to decode")
        if timed:
            self.phase.append(ph_sum)

    def e2e_setup(self):
        self.h2d = int(self.comp_bytes * self.e2e_n)
        self.d2h = self.e2e_n * self.out_bytes
        self.h_out = self.pin(self.e2e_n * self.out_bytes)

    def e2e_step(self):
        _e2e_threads(self.e2e_n, self._e2e_slice)

    def _e2e_slice(self, a, b_):
        descs = self.codecs.decode_batch_host(self.FORMAT, self.host_files[a:b_], self.E2E_ARG, 0,
                                              self.h_out + a * self.out_bytes, self.out_bytes)
        assert all(d.status for d in descs)

    def roofline(self, peak, peak_kind):
        """The dominant kernel (group) of the step: the phase with the largest device time, measured live with CUDA
        events on the launching stream inside the library (gb200_batch_timing). A "launch" is one decode call of
        `sub` images; achieved = algorithmic bytes of that phase for `sub` images / its average duration per call."""
        ph = np.mean(np.array(self.phase), axis=0)
        calls = (self.n + self.sub - 1) // self.sub
        k = max(self.kernel_names, key=lambda q: ph[q])
        per_call_ms = float(ph[k]) / calls
        units = self.n / calls
        alg = self.kernel_bytes[k] * units
        ach = alg / (per_call_ms * 1e-3) / 1e9
        tr, src = traffic_for(self.traffic_keys.get(k, ""), units)
        return {"bound": "hbm", "kernel": self.kernel_names[k], "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                "frac": round(ach / peak, 4), "peak_kind": peak_kind, "traffic": tr, "traffic_source": src,
                "algorithmic_bytes_per_launch": int(alg), "avg_launch_ms": round(per_call_ms, 3),
                "launch": f"one decode call of {units:.0f} images",
                "all_phases": {self.kernel_names[q]: {"ms_per_call": round(float(ph[q]) / calls, 3),
                                                      "GBps": round(self.kernel_bytes[q] * units / (float(ph[q]) / calls * 1e-3) / 1e9, 1),
                                                      "frac": round(self.kernel_bytes[q] * units / (float(ph[q]) / calls * 1e-3) / 1e9 / peak, 4)}
                               for q in self.kernel_names if ph[q] > 0}}

    def extra(self):
        ph = np.mean(np.array(self.phase), axis=0)
        d = {"upload": round(float(ph[0]), 3)}
        for q, name in self.kernel_names.items():
            d[name] = round(float(ph[q]), 3)
        d["host_parse"] = round(float(ph[8]), 3)
        return {"phase_ms_per_step": d, "decode_calls_per_step": (self.n + self.sub - 1) // self.sub}


class JpegWorkload(_BatchDecodeWorkload):
    """BASELINE configs[3]: JPEG baseline decode (Huffman+IDCT+YCbCr), 3840x2160 4:2:0 q90, batch sharded over ranks."""
    name = "JPEG baseline decode (Huffman+IDCT+YCbCr) 3840x2160 4:2:0 q90, batch sharded 1/2/4/8 GPU (BASELINE configs[3])"
    W, H = 3840, 2160
    e2e_api = "gb200_decode_batch_host(JPEG): host file bytes in, rgb8 pixels in pinned host memory out; sub-batches pipelined (download of k overlaps upload + kernels of k+1)"
    FORMAT, E2E_ARG = 0, -1
    kernel_names = {1: "jpeg entropy stage (unstuff/sync/scan/write kernels)", 2: "jpeg_idct_colour_kernel"}
    traffic_keys = {1: "jpeg_entropy", 2: "jpeg_idct_colour_kernel"}
    SUB_BATCH = 512
    E2E_N = 256

    def __init__(self, rank, world, args):
        self._setup(rank, world, args, 4096)
        self.out_bytes = self.W * self.H * 3
        coef = self.W * self.H * 3      # int16 coefficients, 1.5 samples per pixel
        # algorithmic bytes per image: entropy stage = file in + coefficients out; IDCT+colour = coefficients in + rgb8 out
        self.kernel_bytes = {1: self.comp_bytes + coef, 2: coef + self.out_bytes}

    def make_files(self, seed0):
        from PIL import Image as PILImage
        out = []
        for i in range(self.DISTINCT):
            img = synth_photo(self.H, self.W, 3, seed0 + i)
            bio = io.BytesIO()
            PILImage.fromarray(img, "RGB").save(bio, format="JPEG", quality=90, subsampling=2, optimize=False)
            out.append(bio.getvalue())
        return out

    def decode(self, files, dev, stream):
        return self.codecs.jpeg_decode_batch(files, -1, files_dev=dev, stream=stream)

    def config(self):
        return {"units_per_rank": f"{self.n} images {self.W}x{self.H} (total batch {self.total}, {self.DISTINCT} distinct, PIL q90 4:2:0, no DRI)",
                "total_batch": self.total, "sub_batch": self.sub,
                "compressed_bytes_per_image": int(self.comp_bytes), "scaling_note": "strong: total batch fixed, sharded by image index",
                "l2": "inputs larger than L2 (every image has its own device copy)"}

    @staticmethod
    def cpu_run(threads, reps, full):
        from oracle import pyoracle
        W, H = JpegWorkload.W, JpegWorkload.H
        w = JpegWorkload.__new__(JpegWorkload)
        w.DISTINCT = 2
        files = w.make_files(1000)
        nimg = max(threads, 2) if full else 2
        pyoracle.lib()

        def work(t):
            for i in range(t, nimg, threads):
                assert pyoracle.jpeg_load(files[i % len(files)], -1) is not None

        _threads_run(work, threads)
        times = []
        for _ in range(reps):
            t0 = time.perf_counter()
            _threads_run(work, threads)
            times.append(time.perf_counter() - t0)
        return nimg * W * H, times, f"{nimg} images 3840x2160 4:2:0, {threads} thread(s), one image per worker"


class BmpWorkload(_BatchDecodeWorkload):
    """SURVEY 8(f4), first "next" decoder: 24-bit bottom-up BMP (the common case of stbi__bmp_load's easy path),
    3840x2160, decoded to rgb8. Not a BASELINE config: a widening row measured to the same bar."""
    name = "BMP 24-bit decode (BGR swap + vertical flip) 3840x2160, batch 256 (SURVEY 8(f4), not a BASELINE config)"
    W, H = 3840, 2160
    e2e_api = "gb200_decode_batch_host(BMP): host file bytes in, rgb8 pixels in pinned host memory out; sub-batches pipelined"
    FORMAT, E2E_ARG = 7, 0
    kernel_names = {1: "bmp_decode_kernel"}
    traffic_keys = {1: "bmp_decode_kernel"}
    DISTINCT = 4
    E2E_N = 64
    scaling = "weak"

    def __init__(self, rank, world, args):
        self._setup(rank, world, args, 256 * world)
        self.out_bytes = self.W * self.H * 3
        self.kernel_bytes = {1: self.comp_bytes + self.out_bytes}      # the file in, rgb8 out

    def make_files(self, seed0):
        from PIL import Image as PILImage
        out = []
        for i in range(self.DISTINCT):
            bio = io.BytesIO()
            PILImage.fromarray(synth_photo(self.H, self.W, 3, seed0 + i), "RGB").save(bio, format="BMP")
            out.append(bio.getvalue())
        return out

    def decode(self, files, dev, stream):
        return self.codecs.bmp_decode_batch(files, 0, files_dev=dev, stream=stream)

    def config(self):
        return {"units_per_rank": f"{self.n} images {self.W}x{self.H} 24-bit BMP (total batch {self.total}, {self.DISTINCT} distinct)",
                "file_bytes_per_image": int(self.comp_bytes), "scaling_note": "weak: 256 images per GPU",
                "l2": "inputs larger than L2 (every image has its own device copy)"}

    @staticmethod
    def cpu_run(threads, reps, full):
        from oracle import pyoracle
        W, H = BmpWorkload.W, BmpWorkload.H
        w = BmpWorkload.__new__(BmpWorkload)
        w.DISTINCT = 2
        files = w.make_files(1000)
        nimg = max(threads, 2) if full else 2
        pyoracle.lib()

        def work(t):
            for i in range(t, nimg, threads):
                assert pyoracle.bmp_load(files[i % len(files)], 0) is not None

        _threads_run(work, threads)
        times = []
        for _ in range(reps):
            t0 = time.perf_counter()
            _threads_run(work, threads)
            times.append(time.perf_counter() - t0)
        return nimg * W * H, times, f"{nimg} images 3840x2160 24-bit BMP, {threads} thread(s), one image per worker"


class QoixWorkload(_BatchDecodeWorkload):
    """BASELINE configs[4]: QOIX 10-bit LA + LZ4 decode, 2048x2048 images, batch sharded over ranks."""
    name = "QOIX 10-bit LA + LZ4 decode 2048x2048, batch 2048 sharded over ranks (BASELINE configs[4])"
    dtype = "u16"
    W, H = 2048, 2048
    e2e_api = "gb200_decode_batch_host(QOIX): host file bytes in, la16 pixels in pinned host memory out; sub-batches pipelined (download of k overlaps upload + kernels of k+1)"
    FORMAT, E2E_ARG = 3, 0
    E2E_N = 256
    kernel_names = {1: "lz4 kernels (spec/merge/scan/pwrite/parse/resolve)", 2: "qoiplane10 kernels (p10_sync/scan/write/recon)"}
    traffic_keys = {1: "lz4", 2: "qoiplane10"}

    scaling = "weak"                # BASELINE configs[4]: batch 2048 on 8 GPUs = 256 images per GPU at every N

    def __init__(self, rank, world, args):
        self._setup(rank, world, args, 256 * world)
        self.out_bytes = self.W * self.H * 4
        self.kernel_bytes = {1: self.comp_bytes + self.payload, 2: self.payload + self.out_bytes}

    def make_files(self, seed0):
        from oracle import pyoracle            # test-data generation only (the reference's own encoder, restated)
        sys_path_tests()
        from qoixutil import depth_map_la
        out = []
        pay = 0
        for i in range(self.DISTINCT):
            img = depth_map_la(self.H, self.W, seed0 + i, 2)
            f = pyoracle.qoix_encode(img, 10, force_lz4=True)
            pay += int.from_bytes(f[25:29], "big")
            out.append(f)
        self.payload = pay / self.DISTINCT
        return out

    def decode(self, files, dev, stream):
        return self.codecs.qoix_decode_batch(files, 0, files_dev=dev, stream=stream)

    def config(self):
        return {"units_per_rank": f"{self.n} images {self.W}x{self.H} la16 (total batch {self.total}, {self.DISTINCT} distinct, LZ4 forced on)",
                "file_bytes_per_image": int(self.comp_bytes), "opcode_payload_bytes_per_image": int(self.payload),
                "scaling_note": "weak: 256 images per GPU (the batch of 2048 of configs[4] on 8 GPUs), sharded by image index",
                "l2": "inputs larger than L2 (every image has its own device copy)"}

    @staticmethod
    def cpu_run(threads, reps, full):
        from oracle import pyoracle
        W, H = QoixWorkload.W, QoixWorkload.H
        w = QoixWorkload.__new__(QoixWorkload)
        w.DISTINCT = 2
        files = w.make_files(1000)
        nimg = max(threads, 2) if full else 2

        def work(t):
            for i in range(t, nimg, threads):
                assert pyoracle.qoix_decode(files[i % len(files)], 0) is not None

        _threads_run(work, threads)
        times = []
        for _ in range(reps):
            t0 = time.perf_counter()
            _threads_run(work, threads)
            times.append(time.perf_counter() - t0)
        return nimg * W * H, times, f"{nimg} images 2048x2048 la16 + LZ4, {threads} thread(s), one image per worker"


class QoiWorkload(_WorkloadBase):
    """BASELINE configs[0]: QOI decode of one 512x512 RGBA8 image, the plumbing case of examples/convert
    (load with LAYOUT_VERT_STRAIGHT | LAYOUT_GAPLESS; `python -m gamut_b200.convert` is the CLI equivalent).
    Plain QOI has a value-hashed index (qoi.d:536), so one image is one serial chain: this line measures the
    call path, not a throughput kernel. There is no device-resident entry point for QOI: `value` and `e2e` both go
    through gb200_qoi_decode / Image.loadFromMemory with host buffers."""
    name = "QOI decode one 512x512 RGBA8 image via the convert-equivalent path (BASELINE configs[0])"
    dtype = "u8"
    default_steps = 20
    default_e2e_steps = 9
    W = H = 512
    e2e_api = "Image.loadFromMemory(bytes, LAYOUT_VERT_STRAIGHT | LAYOUT_GAPLESS) -> gb200_qoi_decode (host bytes in, malloc'd host pixels out)"

    def __init__(self, rank, world, args):
        import torch
        from gamut_b200 import codecs
        self.torch, self.codecs = torch, codecs
        self.file, self.pixels = make_qoi_file(self.W, self.H)
        self.px_per_step = self.e2e_px_per_step = self.W * self.H
        self.ms = []

    def step(self, stream, timed):
        a = self.torch.cuda.Event(enable_timing=True)
        b = self.torch.cuda.Event(enable_timing=True)
        a.record(stream)
        r = self.codecs.qoi_decode(self.file, 0)
        b.record(stream)
        if r is None:
            raise RuntimeError("qoi_decode failed")
        if not getattr(self, "_checked", False):
            assert np.array_equal(r[0], self.pixels), "QOI decode differs from the encoded pixels"
            self._checked = True

    def config(self):
        return {"units_per_rank": "1 image 512x512 RGBA8 (seeded gradient + noise + flat rectangles + alpha ramp)",
                "file_bytes": len(self.file), "l2": "single small image: L2-resident by nature of the config",
                "note": "value and e2e use the same host-pointer call (no device-resident QOI entry point)"}

    def roofline(self, peak, peak_kind):
        return {"bound": "hbm", "kernel": "qoi_kernel (one thread per image: serial by the format's value-hashed index)",
                "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "peak_kind": peak_kind, "traffic": None,
                "note": "latency-bound plumbing case; no roofline claim"}

    def e2e_setup(self):
        self.h2d = len(self.file)
        self.d2h = self.W * self.H * 4

    def e2e_step(self):
        from gamut_b200.image import Image
        from gamut_b200.types import LAYOUT_VERT_STRAIGHT, LAYOUT_GAPLESS
        im = Image()
        im.loadFromMemory(self.file, LAYOUT_VERT_STRAIGHT | LAYOUT_GAPLESS)
        if im.isError():
            raise RuntimeError(im.errorMessage())

    @staticmethod
    def cpu_run(threads, reps, full):
        from oracle import pyoracle
        f, _ = make_qoi_file(QoiWorkload.W, QoiWorkload.H)
        n = threads if full else 1

        def work(t):
            for _ in range(8):
                assert pyoracle.qoi_decode(f, 0) is not None

        _threads_run(work, n)
        times = []
        for _ in range(reps):
            t0 = time.perf_counter()
            _threads_run(work, n)
            times.append(time.perf_counter() - t0)
        return 8 * n * QoiWorkload.W * QoiWorkload.H, times, f"8 decodes of the 512x512 image per thread, {n} thread(s)"


class QoixEncodeWorkload(_WorkloadBase):
    """SURVEY 8(f1), first encoder row: qoix_lz4_encode's QOI-Plane10 stage (plugins/qoix.d:251, qoiplane10.d:99) on
    the config-5 shape, 2048x2048 10-bit LA. `value`: gb200_qoix_encode_batch_device on device-resident pixels;
    `e2e`: gb200_qoix_encode, host pixels in, malloc'd stream out, image by image."""
    name = "QOI-Plane10 encode 2048x2048 10-bit LA (QOIX encoder, SURVEY 8(f1); shape of BASELINE configs[4])"
    dtype = "u16"
    default_steps = 5
    default_e2e_steps = 3
    W = H = 2048
    N = 128
    DISTINCT = 8
    E2E_IMAGES = 8
    e2e_api = "gb200_qoix_encode (host la16 pixels in, malloc'd QOIX stream out), one call per image"

    def __init__(self, rank, world, args):
        import torch
        from gamut_b200 import codecs
        sys_path_tests()
        from qoixutil import depth_map_la
        self.torch, self.codecs = torch, codecs
        self.n = args.batch or self.N
        self.host = [depth_map_la(self.H, self.W, 100 + rank * 16 + i, 2) for i in range(self.DISTINCT)]
        self.dev = [torch.from_numpy(self.host[i % self.DISTINCT].view(np.int16)).cuda().clone() for i in range(self.n)]
        cap = codecs.qoix_encode_bound(self.W, self.H, 2) + 16
        self.outs = [torch.empty(cap, dtype=torch.uint8, device="cuda") for _ in range(self.n)]
        self.pin_, self.pout = [t.data_ptr() for t in self.dev], [o.data_ptr() for o in self.outs]
        self.shapes = [(self.H, self.W, 2)] * self.n
        self.px_per_step = self.n * self.W * self.H
        self.e2e_px_per_step = self.E2E_IMAGES * self.W * self.H
        self.lens = None
        self.log = EventLog()
        self.kernel_ms = {}

    def step(self, stream, timed):
        e = self.log.span("qe kernels (tile_ne/scan/count/scan/emit)", stream) if timed else None
        self.lens = self.codecs.qoix_encode_batch_device(self.pin_, self.shapes, self.pout, stream.cuda_stream)
        if e:
            e.record(stream)
        if not getattr(self, "_checked", False):
            from oracle import pyoracle          # checker only: the first stream must be the reference encoder's
            exp = pyoracle.qoiplane10_encode(self.host[0])
            got = self.outs[0][:self.lens[0]].cpu().numpy().tobytes()
            assert got == exp, "GPU QOI-Plane10 stream differs from the reference encoder's"
            self._checked = True

    def finish_timing(self):
        self.kernel_ms = self.log.collect()

    def config(self):
        return {"units_per_rank": f"{self.n} images {self.W}x{self.H} la16 ({self.DISTINCT} distinct)",
                "stream_bytes_per_image": int(np.mean(self.lens)) if self.lens else None,
                "l2": "inputs larger than L2 (every image has its own device copy)"}

    def roofline(self, peak, peak_kind):
        avg = {k: float(np.mean(v)) for k, v in self.kernel_ms.items() if v}
        k = max(avg, key=avg.get)
        alg = self.n * (self.W * self.H * 4 + (int(np.mean(self.lens)) if self.lens else 0))
        ach = alg / (avg[k] * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": k, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "peak_kind": peak_kind, "traffic": None, "algorithmic_bytes_per_launch": alg, "avg_launch_ms": round(avg[k], 3),
                "launch": f"one encode call of {self.n} images (pixels are read by three of its five kernels)"}

    def e2e_setup(self):
        self.h2d = self.E2E_IMAGES * self.W * self.H * 4
        self.d2h = self.E2E_IMAGES * (int(np.mean(self.lens)) if self.lens else 0)

    def e2e_step(self):
        for i in range(self.E2E_IMAGES):
            if self.codecs.qoix_encode(self.host[i % self.DISTINCT]) is None:
                raise RuntimeError("qoix_encode failed")

    @staticmethod
    def cpu_run(threads, reps, full):
        from oracle import pyoracle
        sys_path_tests()
        from qoixutil import depth_map_la
        img = depth_map_la(QoixEncodeWorkload.H, QoixEncodeWorkload.W, 100, 2)
        n = threads if full else 1

        def work(t):
            assert pyoracle.qoiplane10_encode(img) is not None

        times = []
        for _ in range(reps):
            t0 = time.perf_counter()
            _threads_run(work, n)
            times.append(time.perf_counter() - t0)
        return n * QoixEncodeWorkload.W * QoixEncodeWorkload.H, times, f"{n} image(s) 2048x2048 la16, {n} thread(s), one image per worker"


def make_qoi_file(w, h):
    """SURVEY 8d cfg 1: seeded gradient + low-amplitude noise + flat rectangles + alpha ramp (tests/qoixutil.py),
    encoded by PIL's QOI writer (an implementation independent of the reference and of this repo).
    Returns (file, pixels)."""
    sys_path_tests()
    from qoixutil import qoi_bytes, qoi_test_image
    px = qoi_test_image(h, w, 4, 1234)
    return qoi_bytes(px), px


def PT_rgba8():
    from gamut_b200.types import PixelType as PT
    return PT.rgba8


def PT_rgbaf32():
    from gamut_b200.types import PixelType as PT
    return PT.rgbaf32


def sys_path_tests():
    import sys
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests")
    if p not in sys.path:
        sys.path.insert(0, p)


WORKLOADS = {"convert": ConvertWorkload, "png": PngWorkload, "jpeg": JpegWorkload, "qoix": QoixWorkload, "qoi": QoiWorkload,
             "bmp": BmpWorkload, "qoix_encode": QoixEncodeWorkload}
