"""One short GPU pass over the TGA codec (one process): the parity tests, then device-resident timings of decode (raw and
run-length files) and encode (CUDA events, 3 warm-up + 5 timed calls). Writes gpurun_out/r2_tga_pytest.txt and
gpurun_out/r2_tga_bench.json.

    gpurun --timeout 200 -- 'timeout 170 python scripts/tga_gpu_pass.py'
"""
import contextlib
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def main():
    import pytest
    t0 = time.time()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        rc = pytest.main(["-x", "-q", "-m", "gpu", "-p", "no:cacheprovider", os.path.join(ROOT, "tests", "test_tga_gpu.py")])
    text = buf.getvalue()
    with open(os.path.join(OUT, "r2_tga_pytest.txt"), "w") as f:
        f.write(text + f"\nexit code {int(rc)}, {time.time() - t0:.1f} s\n")
    print(text[-2500:])
    print("pytest rc", int(rc), flush=True)

    import numpy as np
    import torch
    from gamut_b200 import codecs
    from tgautil import pil_tga
    res = {"pytest_rc": int(rc)}

    def timed(fn, warm=3, reps=5):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in ev:
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)
        return ms[len(ms) // 2], ms

    try:
        rng = np.random.default_rng(0)
        base = []
        for k in range(4):                                       # photo-like: gradient + noise patch + flat areas
            img = np.zeros((1080, 1920, 4), np.uint8)
            img[..., :3] = (np.linspace(0, 255, 1920)[None, :, None] + rng.integers(0, 3, (1080, 1920, 1))).astype(np.uint8)
            img[200:700, 300:1500, :3] = rng.integers(0, 256, (500, 1200, 3))
            img[800:, :, :3] = 40 * k
            img[..., 3] = 255
            base.append(img)
        px = 64 * 1080 * 1920
        for name, rle in (("tga_decode_raw", False), ("tga_decode_rle", True)):
            files = [pil_tga(base[k % 4], rle) for k in range(64)]
            dev = [torch.frombuffer(bytearray(f), dtype=torch.uint8).cuda() for f in files]
            ptrs = [t.data_ptr() for t in dev]
            def run():
                b = codecs.tga_decode_batch(files, files_dev=ptrs)
                b.free()
            med, all_ms = timed(run)
            res[name] = {"workload": "TGA decode, 64 files 1920x1080 32-bit %s, file bytes resident on the device (call includes host header parse and result object)" % ("run-length" if rle else "raw"),
                         "ms": med, "all_ms": all_ms, "Mpx_per_s": px / med / 1e3, "file_bytes": int(sum(len(f) for f in files))}
            del dev
        dev = [torch.from_numpy(base[k % 4]).cuda() for k in range(64)]
        bound = codecs.tga_encode_bound(1920, 1080, 4)
        outs = [torch.empty(bound, dtype=torch.uint8, device="cuda") for _ in range(64)]
        ptrs, optrs, shapes = [t.data_ptr() for t in dev], [o.data_ptr() for o in outs], [(1080, 1920, 4)] * 64
        lens = []
        med, all_ms = timed(lambda: lens.append(codecs.tga_encode_batch_device(ptrs, shapes, optrs)))
        res["tga_encode"] = {"workload": "TGA encode (run-length), 64 rgba8 images 1920x1080, device-resident", "ms": med, "all_ms": all_ms,
                             "Mpx_per_s": px / med / 1e3, "bytes_out": int(sum(lens[-1]))}
    except Exception as e:
        res["bench_error"] = repr(e)
    res["seconds"] = time.time() - t0
    with open(os.path.join(OUT, "r2_tga_bench.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))
    return int(rc)


if __name__ == "__main__":
    sys.exit(main())
