#!/bin/bash
# inflate parity + per-kernel launch durations of the PNG workload (ncu serialises the launches)
timeout 300 python -m pytest tests/test_inflate_gpu.py tests/test_png_gpu.py -m gpu -x -q 2>&1 | tail -4
B=${1:-1024}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/png_launches.csv python bench.py --workload png --batch $B --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/png_ncu.log 2>&1
python - <<'P'
import csv
lines=[l for l in open('gpurun_out/png_launches.csv') if l.startswith('"')]
rows=list(csv.DictReader(lines))
last={}
for x in rows[-13:]:
    print(x['Kernel Name'][:48], round(float(x['Metric Value'].replace(',',''))/1e6,3),'ms')
P
timeout 400 python bench.py --workload png --batch $B --steps 2 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['detail']['phase_ms'])"
