#!/bin/bash
# launch lists (ncu gpu__time_duration) for the PNG and QOIX workloads at bench sizes
T=${1:-r2e}
mkdir -p gpurun_out
for wl in png qoix; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches_$wl.csv python bench.py --workload $wl --only --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${T}_ncu_$wl.log 2>&1
  tail -c 600 gpurun_out/${T}_ncu_$wl.log
done
