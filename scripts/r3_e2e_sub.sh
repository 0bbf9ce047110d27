#!/bin/bash
# e2e of PNG / QOIX for several sub-batch sizes (GB200_E2E_SUB)
for wl in png qoix; do
  for sub in 0 32 64 128 256 1024; do
    GB200_E2E_SUB=$sub timeout 300 python bench.py --workload $wl --only --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$wl sub=$sub value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['step_ms'])"
  done
done
