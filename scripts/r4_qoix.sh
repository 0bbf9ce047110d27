#!/bin/bash
# QOIX state pass: parity tests, bench line, launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_qoix_gpu.py -x -q -m gpu > gpurun_out/r4_pytest_qoix.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r4_pytest_qoix.txt
timeout 600 python bench.py --workload qoix --steps 5 --warmup 3 > gpurun_out/r4_bench_qoix.json 2> gpurun_out/r4_bench_qoix.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r4_bench_qoix.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline'].get('all_phases'), d['e2e']['value'])
P
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4_launches_qoix.csv python bench.py --workload qoix --only --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r4_launches_qoix.csv | head -30
# full ncu captures (with source) of the QOI-Plane10 kernels
for K in ; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:$K -s 1 -c 1 -f -o gpurun_out/r4_prof_$K python bench.py --workload qoix --only --batch 148 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r4_prof_$K.log 2>&1
ls -la gpurun_out/r4_prof_$K.ncu-rep
done
