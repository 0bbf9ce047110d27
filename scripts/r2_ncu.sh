#!/bin/bash
# usage: r2_ncu.sh TAG WORKLOAD KERNEL_REGEX BATCH [SKIP]   -- one ncu --set full capture on a small batch
T=$1; WL=$2; K=$3; B=$4; S=${5:-0}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c 2 -o gpurun_out/${T}_prof -f \
  python bench.py --workload $WL --only --batch $B --sub-batch $B --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${T}_ncu_full.log 2>&1
tail -c 300 gpurun_out/${T}_ncu_full.log; ls -la gpurun_out/${T}_prof.ncu-rep
