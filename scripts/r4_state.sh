#!/bin/bash
# state pass: parity suite, smoke, the default bench line with all workloads, reference arm, launch lists
T=${1:-r4}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${T}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1
( time timeout 1500 python bench.py ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>&1
for wl in png qoix; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches_$wl.csv python bench.py --workload $wl --only --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${T}_ncu_$wl.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_jpeg.csv python bench.py --workload jpeg --only --batch 512 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${T}_ncu_jpeg.log 2>&1
cat gpurun_out/${T}_pytest_gpu.txt; tail -2 gpurun_out/${T}_smoke.txt
cut -c1-1200 gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
cut -c1-400 gpurun_out/${T}_bench_reference.json
