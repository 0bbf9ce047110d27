#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_qoix_encode_gpu.py -x -q -m gpu > gpurun_out/r4_pytest_encode.txt 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r4_pytest_encode.txt
timeout 300 python scripts/qoix_encode_bench.py 256 > gpurun_out/r4_encode_bench.json 2> gpurun_out/r4_encode_bench.err; cat gpurun_out/r4_encode_bench.json; tail -3 gpurun_out/r4_encode_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:qe_ -c 10 --csv --log-file gpurun_out/r4_launches_encode.csv python scripts/qoix_encode_bench.py 64 > /dev/null 2>&1
python - <<'P'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r4_launches_encode.csv') if l.startswith('"')))
h=rows[0]; ki,mi,vi=h.index('Kernel Name'),h.index('Metric Name'),h.index('Metric Value')
for r in rows[1:16]: print(r[ki][:40], r[mi][:30], r[vi])
P
