#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_inflate_gpu.py tests/test_png_gpu.py -x -q -m gpu > gpurun_out/r4_pytest_png.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r4_pytest_png.txt
timeout 600 python bench.py --workload png --only --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r4_bench_png.json 2> gpurun_out/r4_bench_png.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r4_bench_png.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline'].get('avg_launch_ms'), d.get('detail',{}).get('phase_ms_per_step') or list(d.get('detail',{}).items())[:6])
P
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r4_launches_png.csv python bench.py --workload png --only --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r4_launches_png.csv | grep "infp\|unfilter_kernel\|gather_seg" | cut -c1-150
