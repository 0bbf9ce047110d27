#!/usr/bin/env python
"""DRAM traffic per kernel (group) of THIS build -> profiles/traffic.json (read by benchlib.traffic_for).

Runs on the GPU box (under gpurun): one `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` pass per workload on a
reduced batch (ncu replays every launch; the figures are per unit -- per image, per 8192^2 convert launch -- and bench.py
scales them to the units of its own launches), groups the launches by kernel, and writes bytes per unit:

    gpurun -- 'python scripts/capture_traffic.py'      # writes gpurun_out/traffic.json; copy it to profiles/traffic.json

Keys: convert_direct<rgba8,rgbaf32>, convert_direct<rgbaf32,rgba8>, inflate, unfilter, jpeg_entropy,
jpeg_idct_colour_kernel, lz4, qoiplane10."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")

GROUPS = {   # workload -> (bench args, units per step, [(key, kernel-name regex)])
    "convert": (["--workload", "convert", "--steps", "2"], 1,
                [("convert_direct<rgba8,rgbaf32>", r"convert_direct.*<\(?(int\))?12, \(?(int\))?14"),
                 ("convert_direct<rgbaf32,rgba8>", r"convert_direct.*<\(?(int\))?14, \(?(int\))?12")]),
    "jpeg": (["--workload", "jpeg", "--batch", "64", "--sub-batch", "64", "--steps", "1"], 64,
             [("jpeg_entropy", r"jpeg_(unstuff_count|unstuff_scan|unstuff_write|sync|repair|scan|write|huffman)_kernel"), ("jpeg_idct_colour_kernel", r"jpeg_idct_colour_kernel")]),
    "png": (["--workload", "png", "--batch", "64", "--steps", "1"], 64,
            [("inflate", r"infp_|inflate_batch_kernel|gather_segments"), ("unfilter", r"unfilter_kernel")]),
    "qoix": (["--workload", "qoix", "--batch", "32", "--steps", "1"], 32,
             [("lz4", r"lz4_"), ("qoiplane10", r"p10_")]),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    result = {}
    for wl, (args, units, groups) in GROUPS.items():
        log = os.path.join(OUT, "traffic_%s.csv" % wl)
        cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--csv", "--log-file", log,
               sys.executable, os.path.join(ROOT, "bench.py"), "--only", "--warmup", "3", "--no-cpu-baseline", "--e2e-steps", "0"] + args
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
        if r.returncode != 0:
            result[wl + "_error"] = (r.stdout + r.stderr)[-400:]
            continue
        rows = list(csv.reader(l for l in open(log) if l.startswith('"')))
        hdr = rows[0]
        ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        per_kernel = collections.defaultdict(float)
        launches = collections.Counter()
        for row in rows[1:]:
            v = float(row[vi].replace(",", ""))
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(row[ui], 1)
            per_kernel[row[ki]] += v
            if row[mi] == "dram__bytes_read.sum":
                launches[row[ki]] += 1
        nsteps = 3 + int(args[args.index("--steps") + 1])          # warm-up steps run the same launches
        # the PNG workload also runs unfilter-only legs (after the timed region): their launches are part of the kernel
        # totals, so "unfilter" is reported per launch instead of per step
        for key, pat in groups:
            tot = sum(v for k, v in per_kernel.items() if re.search(pat, k))
            nl = sum(n for k, n in launches.items() if re.search(pat, k))
            if tot == 0:
                continue
            per_unit = tot / nl / units if key == "unfilter" else tot / (nsteps * units)
            result[key] = {"bytes_per_unit": per_unit, "unit": "8192x8192 launch" if wl == "convert" else "image",
                           "source": "scripts/capture_traffic.py: ncu dram__bytes_read.sum + dram__bytes_write.sum, %s, %d units per launch"
                                     % (" ".join(args), units)}
    json.dump(result, open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
    print(json.dumps(result, indent=1)[:3000])


if __name__ == "__main__":
    main()
