#!/bin/bash
for t in 1 2 3; do
  echo "png threads=$t"; GB200_E2E_THREADS=$t timeout 300 python bench.py --workload png --only --steps 1 --e2e-steps 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['e2e']['value'], d['e2e']['step_ms'])"
done
for t in 1 2; do
  echo "qoix threads=$t sub=64"; GB200_E2E_SUB=64 GB200_E2E_THREADS=$t timeout 300 python bench.py --workload qoix --only --steps 1 --e2e-steps 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['e2e']['value'], d['e2e']['step_ms'])"
done
