#!/bin/bash
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 40 --csv --log-file gpurun_out/qoix_launches.csv python bench.py --workload qoix --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/qoix_ncu.log 2>&1
python - <<'P'
import csv
lines=[l for l in open('gpurun_out/qoix_launches.csv') if l.startswith('"')]
rows=list(csv.DictReader(lines))
for x in rows[-24:]:
    print(x['Kernel Name'][:48], x['Metric Name'][:24], x['Metric Value'])
P
