#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_qoix_gpu.py tests/test_image_gpu.py tests/test_qoix_encode_gpu.py -x -q -m gpu > gpurun_out/r4_pytest_qoi.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r4_pytest_qoi.txt
timeout 300 python bench.py --workload qoi --only > gpurun_out/r4_bench_qoi.json 2> gpurun_out/r4_bench_qoi.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r4_bench_qoi.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'])
P
