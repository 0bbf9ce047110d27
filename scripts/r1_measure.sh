#!/bin/bash
# round-1 measurement pass: GPU tests, bench lines for every workload, ncu launch lists
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1_pytest_gpu.txt
for wl in convert png jpeg qoix; do
  timeout 600 python bench.py --workload $wl > gpurun_out/r1_bench_$wl.json 2> gpurun_out/r1_bench_$wl.err
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference.json 2>&1
for wl in png jpeg qoix; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_$wl.csv python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r1_ncu_$wl.log 2>&1
done
cat gpurun_out/r1_pytest_gpu.txt
for wl in convert png jpeg qoix; do cut -c1-1500 gpurun_out/r1_bench_$wl.json; tail -3 gpurun_out/r1_bench_$wl.err; done
