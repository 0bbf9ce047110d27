#!/bin/bash
# PNG / QOIX iteration pass: parity, bench lines, launch lists
T=${1:-r2f}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_pytest_gpu.txt
for wl in png qoix; do
  timeout 900 python bench.py --workload $wl --only --no-cpu-baseline > gpurun_out/${T}_bench_$wl.json 2> gpurun_out/${T}_bench_$wl.err
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches_$wl.csv python bench.py --workload $wl --only --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${T}_ncu_$wl.log 2>&1
done
cat gpurun_out/${T}_pytest_gpu.txt
for wl in png qoix; do cut -c1-1500 gpurun_out/${T}_bench_$wl.json; tail -3 gpurun_out/${T}_bench_$wl.err; done
