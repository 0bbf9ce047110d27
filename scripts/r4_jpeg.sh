#!/bin/bash
# JPEG state pass: parity tests, bench line (batch 1024 to keep it short), launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_jpeg_gpu.py tests/test_image_gpu.py tests/test_batch_host_gpu.py -x -q -m gpu > gpurun_out/r4_pytest_jpeg.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r4_pytest_jpeg.txt
timeout 600 python bench.py --workload jpeg --only --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r4_bench_jpeg.json 2> gpurun_out/r4_bench_jpeg.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r4_bench_jpeg.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline'].get('all_phases'), d.get('detail',{}).get('phase_ms_per_step'))
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r4_launches_jpeg.csv python bench.py --workload jpeg --only --batch 512 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r4_launches_jpeg.csv | grep -v "fill\|read_back" | head -20
if [ -n "$1" ]; then
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"$1" -c 2 -f -o gpurun_out/r4_prof_jpeg python bench.py --workload jpeg --only --batch 512 --sub-batch 512 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r4_prof_jpeg.log 2>&1
ls -la gpurun_out/r4_prof_jpeg.ncu-rep
fi
