#!/bin/bash
# quick end-to-end leg check of all four workloads (small batches)
timeout 300 python -m pytest tests/test_convert_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('convert', d['value'], d['e2e'])"
for w in png jpeg qoix; do
timeout 400 python bench.py --workload $w --batch 128 --steps 1 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', d['value'], d['e2e']['value'], d['detail']['phase_ms'])"
done
