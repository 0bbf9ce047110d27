#!/bin/bash
# round-1 measurement pass: GPU tests, smoke, bench lines for every workload, reference arm, ncu launch lists
T=${1:-r1f}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${T}_pytest_gpu.txt
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.txt 2>&1
for wl in convert png jpeg qoix; do
  timeout 900 python bench.py --workload $wl > gpurun_out/${T}_bench_$wl.json 2> gpurun_out/${T}_bench_$wl.err
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>&1
for wl in png jpeg qoix; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_$wl.csv python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${T}_ncu_$wl.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_convert.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${T}_ncu_convert.log 2>&1
cat gpurun_out/${T}_pytest_gpu.txt; tail -2 gpurun_out/${T}_smoke.txt
for wl in convert png jpeg qoix; do cut -c1-1200 gpurun_out/${T}_bench_$wl.json; tail -3 gpurun_out/${T}_bench_$wl.err; done
cut -c1-600 gpurun_out/${T}_bench_reference.json
