#!/bin/bash
# compute-sanitizer memcheck over the small GPU parity tests (slow: 10-50x); results under gpurun_out/
mkdir -p gpurun_out
export GB200_SANITIZE=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 --log-file gpurun_out/sanitize_memcheck.log \
  python -m pytest tests/test_inflate_gpu.py tests/test_jpeg_gpu.py tests/test_png_gpu.py tests/test_qoix_gpu.py tests/test_convert_gpu.py -m gpu -x -q \
  -k "not config4 and not 4k and not config5 and not long_blocks and not config3 and not geometry" > gpurun_out/sanitize_pytest.txt 2>&1
echo "exit $?" >> gpurun_out/sanitize_pytest.txt
tail -5 gpurun_out/sanitize_pytest.txt
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/sanitize_memcheck.log
grep -m 12 -A6 "Invalid\|misaligned" gpurun_out/sanitize_memcheck.log | head -80
tail -3 gpurun_out/sanitize_memcheck.log
