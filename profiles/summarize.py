#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.
usage: summarize.py full <rep.ncu-rep> <out.txt> | launches <launches.csv> <out.txt>"""
import csv
import subprocess
import sys
from collections import defaultdict

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
        "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "smsp__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
        "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct"]


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of {rep}\n")
        for r in rows[2:]:
            f.write("\nkernel: " + r[hdr.index("Kernel Name")][:160] + "\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"  {w:90s} {r[i]} {units[i]}\n")


def launches(path, out):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(list)
    for r in rows[1:]:
        agg[r[k]].append(float(r[v].replace(",", "")))
    tot = sum(sum(x) for x in agg.values())
    with open(out, "w") as f:
        f.write(f"# launch list summary of {path} (gpu__time_duration.sum, ns; cold-cache, serialised: compare shares)\n")
        for name, x in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{sum(x) / tot * 100:6.2f}%  n={len(x):4d}  avg={sum(x) / len(x) / 1e3:10.1f} us  {name[:140]}\n")


if __name__ == "__main__":
    {"full": full, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
