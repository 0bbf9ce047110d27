#!/bin/bash
# full ncu capture (with source) of one kernel of a workload: $1 = kernel regex, $2 = tag, $3 = batch, $4 = workload
K=${1:-infp_resolve_kernel}; T=${2:-resolve}; B=${3:-296}; W=${4:-png}
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:$K -c 1 -f -o gpurun_out/prof_$T python bench.py --workload $W --only --batch $B --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/prof_$T.log 2>&1
tail -2 gpurun_out/prof_$T.log | cut -c1-300
ls -la gpurun_out/prof_$T.ncu-rep
