#!/bin/bash
mkdir -p gpurun_out
export GB200_SANITIZE=1
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 --log-file gpurun_out/r4_sanitize2_memcheck.log \
  python -m pytest tests/test_jpeg_gpu.py tests/test_batch_host_gpu.py tests/test_qoix_gpu.py tests/test_inflate_gpu.py -m gpu -x -q \
  -k "not config4 and not 4k and not config5 and not long_blocks" > gpurun_out/r4_sanitize2_pytest.txt 2>&1
echo "exit $?" >> gpurun_out/r4_sanitize2_pytest.txt
tail -4 gpurun_out/r4_sanitize2_pytest.txt
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r4_sanitize2_memcheck.log
grep -m 8 -A8 "Invalid\|misaligned" gpurun_out/r4_sanitize2_memcheck.log | head -60
tail -2 gpurun_out/r4_sanitize2_memcheck.log
