#!/bin/bash
# full ncu capture (with source) of one kernel of the PNG workload: $1 = kernel regex, $2 = tag, $3 = batch
K=${1:-infp_resolve_kernel}; T=${2:-resolve}; B=${3:-296}
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:$K -c 1 -f -o gpurun_out/prof_$T python bench.py --workload png --only --batch $B --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/prof_$T.log 2>&1
tail -3 gpurun_out/prof_$T.log
ls -la gpurun_out/prof_$T.ncu-rep
timeout 120 python scripts/hw_decompress_bar.py > gpurun_out/hw_bar.json 2> gpurun_out/hw_bar.err; cat gpurun_out/hw_bar.json; tail -3 gpurun_out/hw_bar.err
