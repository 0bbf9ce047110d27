#!/bin/bash
# round 2, first GPU pass: parity suite, the new default bench line (JPEG 4096 + detail.workloads), reference arm
T=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/${T}_gpu.txt; nproc >> gpurun_out/${T}_gpu.txt; free -g >> gpurun_out/${T}_gpu.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${T}_pytest_gpu.txt
( time timeout 1200 python bench.py ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>&1
cat gpurun_out/${T}_pytest_gpu.txt
cut -c1-3000 gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
cut -c1-600 gpurun_out/${T}_bench_reference.json
