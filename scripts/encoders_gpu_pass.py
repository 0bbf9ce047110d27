"""One short GPU pass over the encoders (one process, so the box pays for one torch import): the parity tests of the
QOI / QOIX (all four sub-codecs) / BMP encoders, then device-resident timings of the new ones (CUDA events, inputs larger
than L2, 3 warm-up + 5 timed launches). Writes gpurun_out/r2_encoders_pytest.txt and gpurun_out/r2_encoders_bench.json.

    gpurun --timeout 200 -- 'timeout 170 python scripts/encoders_gpu_pass.py'
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
        rc = pytest.main(["-q", "-m", "gpu", "-p", "no:cacheprovider", "-rxX"] + [os.path.join(ROOT, "tests", f) for f in (
            "test_qoi_encode_gpu.py", "test_qoix_encode_gpu.py", "test_zz_qoi2avg_encode_gpu.py", "test_zz_qoi10b_encode_gpu.py",
            "test_zz_bmp_encode_gpu.py")])
    text = buf.getvalue()
    with open(os.path.join(OUT, "r2_encoders_pytest.txt"), "w") as f:
        f.write(text + f"\nexit code {int(rc)}, {time.time() - t0:.1f} s\n")
    print(text[-3000:])
    print("pytest rc", int(rc), flush=True)

    import numpy as np
    import torch
    from gamut_b200 import codecs
    from qoixutil import depth_map_la, qoi_test_image
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
        # QOI: 64 rgba8 images 1920x1080 (8 distinct), 531 MB of pixels
        base = [qoi_test_image(1080, 1920, 4, 100 + k) for k in range(8)]
        dev = [torch.from_numpy(base[k % 8]).cuda() for k in range(64)]
        bound = codecs.qoi_encode_bound(1920, 1080, 4) + 16
        outs = [torch.empty(bound, dtype=torch.uint8, device="cuda") for _ in range(64)]
        ptrs, optrs, shapes = [t.data_ptr() for t in dev], [o.data_ptr() for o in outs], [(1080, 1920, 4)] * 64
        lens = []
        med, all_ms = timed(lambda: lens.append(codecs.qoi_encode_batch_device(ptrs, shapes, optrs)))
        px = 64 * 1080 * 1920
        res["qoi_encode"] = {"workload": "QOI encode, 64 rgba8 images 1920x1080, device-resident (call includes the length read-back)",
                             "ms": med, "all_ms": all_ms, "Mpx_per_s": px / med / 1e3, "bytes_out": int(sum(lens[-1])),
                             "GBps_in_plus_out": (px * 4 + sum(lens[-1])) / med / 1e6}
        del dev, outs
        # QOI-Plane: 64 la8 images 2048x2048 (4 distinct)
        base = [(depth_map_la(2048, 2048, 200 + k, 2) >> 8).astype(np.uint8) for k in range(4)]
        dev = [torch.from_numpy(base[k % 4]).cuda() for k in range(64)]
        bound = codecs.qoix_encode_bound(2048, 2048, 2) + 16
        outs = [torch.empty(bound, dtype=torch.uint8, device="cuda") for _ in range(64)]
        ptrs, optrs, shapes = [t.data_ptr() for t in dev], [o.data_ptr() for o in outs], [(2048, 2048, 2)] * 64
        lens = []
        med, all_ms = timed(lambda: lens.append(codecs.qoix_encode_batch_device(ptrs, shapes, optrs, bitdepths=[8] * 64)))
        px = 64 * 2048 * 2048
        res["qoiplane_encode"] = {"workload": "QOI-Plane encode, 64 la8 images 2048x2048, device-resident",
                                  "ms": med, "all_ms": all_ms, "Mpx_per_s": px / med / 1e3, "bytes_out": int(sum(lens[-1])),
                                  "GBps_in_plus_out": (px * 2 + sum(lens[-1])) / med / 1e6}
        del dev, outs
        # QOI2AVG (8-bit) and QOI-10b (10-bit): 64 rgba images 1920x1080 through the same entry point
        for name, imgs, bd in (("qoi2avg_encode", [qoi_test_image(1080, 1920, 4, 300 + k) for k in range(4)], 8),
                               ("qoi10b_encode", [(qoi_test_image(1080, 1920, 4, 400 + k).astype(np.uint16) * 257) & 0xffc0 for k in range(4)], 10)):
            dev = [torch.from_numpy(imgs[k % 4].view(np.int16) if bd == 10 else imgs[k % 4]).cuda() for k in range(64)]
            outs = [torch.empty(1080 * 1920 * 7 + 256, dtype=torch.uint8, device="cuda") for _ in range(64)]
            ptrs, optrs, shapes = [t.data_ptr() for t in dev], [o.data_ptr() for o in outs], [(1080, 1920, 4)] * 64
            lens = []
            med, all_ms = timed(lambda: lens.append(codecs.qoix_encode_batch_device(ptrs, shapes, optrs, bitdepths=[bd] * 64)))
            px = 64 * 1080 * 1920
            res[name] = {"workload": "%s, 64 rgba images 1920x1080, device-resident" % name, "ms": med, "all_ms": all_ms,
                         "Mpx_per_s": px / med / 1e3, "bytes_out": int(sum(lens[-1]))}
            del dev, outs
    except Exception as e:                                      # the parity result above is what matters most
        res["bench_error"] = repr(e)
    res["seconds"] = time.time() - t0
    with open(os.path.join(OUT, "r2_encoders_bench.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))
    return int(rc)


if __name__ == "__main__":
    sys.exit(main())
