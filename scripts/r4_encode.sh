#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_qoix_encode_gpu.py -x -q -m gpu > gpurun_out/r4_pytest_encode.txt 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r4_pytest_encode.txt
timeout 300 python scripts/qoix_encode_bench.py 256 > gpurun_out/r4_encode_bench.json 2> gpurun_out/r4_encode_bench.err; cat gpurun_out/r4_encode_bench.json; tail -3 gpurun_out/r4_encode_bench.err
