#!/usr/bin/env python
"""Per-source-line totals of an ncu `--page source --csv` SASS export, using nvdisasm -g line info.
usage: hotlines.py <ncu_source.csv> <nvdisasm_-g_-c.sass> <mangled-kernel-substring> [topN]"""
import csv, re, sys
from collections import defaultdict

src_csv, sass, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# ---- offset -> (file, line) for the kernel
lines = open(sass).read().split("\n")
inside = False
cur = ("?", 0)
off2line = {}
for l in lines:
    if l.startswith(".text."):
        inside = kname in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, ii, isamp, ithr = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
base = None
agg = defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in rows[2:]:
    if not r or not r[ia].startswith('0x'):
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    k = off2line.get(a - base, ("?", 0))
    v = (int(r[ii]), int(r[isamp]), int(r[ithr]))
    for q in range(3):
        agg[k][q] += v[q]; tot[q] += v[q]
print("total inst %d samples %d thread-inst %d (avg active lanes %.1f)" % (tot[0], tot[1], tot[2], tot[2] / max(tot[0], 1)))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-22s %5d  inst %5.1f%%  samples %5.1f%%  lanes %4.1f" % (k[0], k[1], 100 * v[0] / tot[0], 100 * v[1] / max(tot[1], 1), v[2] / max(v[0], 1)))
