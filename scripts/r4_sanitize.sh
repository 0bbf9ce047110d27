#!/bin/bash
# compute-sanitizer memcheck over the GPU parity tests of the kernels rewritten in the final pass of round 2
mkdir -p gpurun_out
export GB200_SANITIZE=1
timeout 1700 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 --log-file gpurun_out/r4_sanitize_memcheck.log \
  python -m pytest tests/test_jpeg_gpu.py tests/test_qoix_gpu.py tests/test_qoix_encode_gpu.py -m gpu -x -q \
  -k "not config4 and not 4k and not config5" > gpurun_out/r4_sanitize_pytest.txt 2>&1
echo "exit $?" >> gpurun_out/r4_sanitize_pytest.txt
tail -5 gpurun_out/r4_sanitize_pytest.txt
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r4_sanitize_memcheck.log
grep -m 12 -A8 "Invalid\|misaligned" gpurun_out/r4_sanitize_memcheck.log | head -90
tail -3 gpurun_out/r4_sanitize_memcheck.log
