#!/usr/bin/env python
"""A stated bar, not a product path: times the chip's own decompression engine (cuMemBatchDecompressAsync,
cuda.h: CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE / _LZ4) on the same streams the PNG and QOIX workloads of bench.py feed to
the hand-written pipelines. Writes one JSON object to stdout:

  {"deflate": {...}, "lz4": {...}, "device": {"algorithm_mask": m, "max_length": L}}

Each leg: n streams, compressed / raw bytes per stream, best-of-5 ms, GB/s of output, Mpixels/s equivalent, and whether
the engine's output equals zlib's / the oracle's (it must, or the number is not a bar). When the engine refuses the
batch (unsupported algorithm, stream longer than the device limit) the leg says why instead.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Params(ctypes.Structure):          # CUmemDecompressParams, cuda.h
    _fields_ = [("srcNumBytes", ctypes.c_size_t), ("dstNumBytes", ctypes.c_size_t), ("dstActBytes", ctypes.c_void_p),
                ("src", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("algo", ctypes.c_int), ("padding", ctypes.c_ubyte * 20)]


ALG_DEFLATE, ALG_LZ4 = 1, 4
ATTR_MASK, ATTR_MAXLEN = 136, 137


def main():
    import torch
    import benchlib

    torch.cuda.init()
    torch.zeros(1, device="cuda")                     # primary context, current on this thread
    cu = ctypes.CDLL("libcuda.so.1")

    def ck(e, what):
        if e != 0:
            raise RuntimeError(f"{what}: CUresult {e}")

    dev = ctypes.c_int()
    ck(cu.cuCtxGetDevice(ctypes.byref(dev)), "cuCtxGetDevice")
    v = ctypes.c_int()
    mask = maxlen = 0
    if cu.cuDeviceGetAttribute(ctypes.byref(v), ATTR_MASK, dev) == 0:
        mask = v.value
    if cu.cuDeviceGetAttribute(ctypes.byref(v), ATTR_MAXLEN, dev) == 0:
        maxlen = v.value
    out = {"device": {"algorithm_mask": int(mask), "max_length": int(maxlen),
                      "deflate": bool(mask & 1), "snappy": bool(mask & 2), "lz4": bool(mask & 4)}}
    if not hasattr(cu, "cuMemBatchDecompressAsync"):
        out["unavailable"] = "libcuda has no cuMemBatchDecompressAsync"
        print(json.dumps(out))
        return

    def run(name, algo, streams, raws, n, px_per_stream):
        """streams: distinct compressed byte strings; raws: their expected outputs; n: batch size (streams cycle)."""
        leg = {"streams": n, "distinct": len(streams), "compressed_bytes_per_stream": int(np.mean([len(s) for s in streams])),
               "raw_bytes_per_stream": int(np.mean([len(r) for r in raws]))}
        out[name] = leg
        if not (mask & int(algo)):
            leg["unavailable"] = "algorithm not in CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_ALGORITHM_MASK"
            return
        if maxlen and max(len(r) for r in raws) > maxlen:
            leg["note"] = f"raw stream ({max(len(r) for r in raws)} B) exceeds MEM_DECOMPRESS_MAXIMUM_LENGTH ({maxlen} B); tried anyway"
        nd = len(streams)
        cstride = (max(len(s) for s in streams) + 255) // 256 * 256
        rstride = (max(len(r) for r in raws) + 255) // 256 * 256

        def alloc(nbytes):                            # plain cuMemAlloc: decompress-capable by definition
            p = ctypes.c_uint64()
            ck(cu.cuMemAlloc_v2(ctypes.byref(p), ctypes.c_size_t(nbytes)), "cuMemAlloc")
            return p.value
        d_src, d_dst, d_act = alloc(cstride * n), alloc(rstride * n), alloc(4 * n)
        try:
            hsrc = np.zeros((nd, cstride), np.uint8)
            for i, s in enumerate(streams):
                hsrc[i, :len(s)] = np.frombuffer(s, np.uint8)
            for i in range(n):
                ck(cu.cuMemcpyHtoD_v2(ctypes.c_uint64(d_src + i * cstride), ctypes.c_void_p(hsrc[i % nd].ctypes.data), ctypes.c_size_t(cstride)), "HtoD")
            params = (_Params * n)()
            for i in range(n):
                p = params[i]
                p.srcNumBytes = len(streams[i % nd]); p.dstNumBytes = len(raws[i % nd])
                p.dstActBytes = d_act + 4 * i; p.src = d_src + i * cstride; p.dst = d_dst + i * rstride; p.algo = int(algo)
            stream = ctypes.c_void_p()
            ck(cu.cuStreamCreate(ctypes.byref(stream), 0), "cuStreamCreate")
            ev0, ev1 = ctypes.c_void_p(), ctypes.c_void_p()
            ck(cu.cuEventCreate(ctypes.byref(ev0), 0), "event"); ck(cu.cuEventCreate(ctypes.byref(ev1), 0), "event")
            times = []
            for rep in range(6):
                ck(cu.cuMemsetD8Async(ctypes.c_uint64(d_dst), 0, ctypes.c_size_t(rstride * n), stream), "memset")
                ck(cu.cuEventRecord(ev0, stream), "record")
                erri = ctypes.c_size_t(0)
                e = cu.cuMemBatchDecompressAsync(params, ctypes.c_size_t(n), 0, ctypes.byref(erri), stream)
                if e != 0:
                    leg["unavailable"] = f"cuMemBatchDecompressAsync: CUresult {e} (errorIndex {erri.value})"
                    return
                ck(cu.cuEventRecord(ev1, stream), "record")
                e = cu.cuStreamSynchronize(stream)
                if e != 0:
                    leg["unavailable"] = f"stream sync after decompress: CUresult {e}"
                    return
                ms = ctypes.c_float()
                ck(cu.cuEventElapsedTime(ctypes.byref(ms), ev0, ev1), "elapsed")
                if rep:
                    times.append(ms.value)
            okall = True
            for i in range(min(nd, n)):
                h = np.empty(len(raws[i]), np.uint8)
                ck(cu.cuMemcpyDtoH_v2(ctypes.c_void_p(h.ctypes.data), ctypes.c_uint64(d_dst + i * rstride), ctypes.c_size_t(len(raws[i]))), "DtoH")
                okall &= bool(np.array_equal(h, np.frombuffer(raws[i], np.uint8)))
            act = np.empty(n, np.uint32)
            ck(cu.cuMemcpyDtoH_v2(ctypes.c_void_p(act.ctypes.data), ctypes.c_uint64(d_act), ctypes.c_size_t(4 * n)), "DtoH")
            best = min(times)
            total_raw = sum(len(raws[i % nd]) for i in range(n))
            leg.update({"ms_best_of_5": round(best, 3), "ms_all": [round(t, 3) for t in times],
                        "out_GBps": round(total_raw / (best * 1e-3) / 1e9, 1),
                        "Mpixels_s_equivalent": round(n * px_per_stream / (best * 1e-3) / 1e6, 1),
                        "output_matches": okall,
                        "dstActBytes_ok": bool((act[:min(nd, n)] == np.array([len(r) for r in raws[:min(nd, n)]], np.uint32)).all())})
        finally:
            for p in (d_src, d_dst, d_act):
                cu.cuMemFree_v2(ctypes.c_uint64(p))

    # ---- DEFLATE: the zlib streams of the PNG workload (configs[2]), without the 2-byte zlib header and the Adler-32
    W, H = benchlib.PngWorkload.W, benchlib.PngWorkload.H
    files = benchlib.make_png_files(benchlib.PngWorkload.DISTINCT, W, H, 1000)
    z = [benchlib.split_idat(f) for f in files]
    raws = [zlib.decompress(s) for s in z]
    n_png = int(os.environ.get("HWBAR_PNG_N", "1024"))
    run("deflate", ALG_DEFLATE, [s[2:-4] for s in z], raws, n_png, W * H)

    # ---- LZ4: the LZ4 block of the QOIX workload's files (configs[4]); payload = header(25) + u32 size + block
    from oracle import pyoracle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from qoixutil import depth_map_la
    QW, QH = benchlib.QoixWorkload.W, benchlib.QoixWorkload.H
    blocks, plain = [], []
    lz4 = ctypes.CDLL(ctypes.util.find_library("lz4") or "liblz4.so.1")
    for i in range(4):
        f = pyoracle.qoix_encode(depth_map_la(QH, QW, 1000 + i, 2), 10, force_lz4=True)
        orig = int.from_bytes(f[25:29], "big")
        blk = f[29:]
        dst = ctypes.create_string_buffer(orig)
        got = lz4.LZ4_decompress_safe(blk, dst, len(blk), orig)
        if got != orig:
            out["lz4"] = {"unavailable": f"could not cut the LZ4 block out of the QOIX file (liblz4 says {got}, header says {orig})"}
            break
        blocks.append(blk)
        plain.append(dst.raw)
    else:
        run("lz4", ALG_LZ4, blocks, plain, int(os.environ.get("HWBAR_QOIX_N", "256")), QW * QH)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
