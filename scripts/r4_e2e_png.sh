#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_batch_host_gpu.py tests/test_png_gpu.py -x -q -m gpu 2>&1 | tail -3
for p in 1 2 3 4; do
  echo "png parts=$p"; GB200_E2E_PARTS=$p timeout 300 python bench.py --workload png --only --steps 1 --e2e-steps 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['e2e']['value'], d['e2e']['step_ms'])"
done
echo "png default"; timeout 300 python bench.py --workload png --only --steps 1 --e2e-steps 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['e2e']['value'], d['e2e']['step_ms'])"
