#!/bin/bash
# JPEG iteration pass: parity, bench (JPEG only), launch list
T=${1:-r2b}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_pytest_gpu.txt
timeout 900 python bench.py --workload jpeg --only --no-cpu-baseline > gpurun_out/${T}_bench_jpeg.json 2> gpurun_out/${T}_bench_jpeg.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_jpeg.csv python bench.py --workload jpeg --only --batch 512 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${T}_ncu_jpeg.log 2>&1
cat gpurun_out/${T}_pytest_gpu.txt
cut -c1-2500 gpurun_out/${T}_bench_jpeg.json; tail -5 gpurun_out/${T}_bench_jpeg.err
