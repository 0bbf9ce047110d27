#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over small cases of the kernels rewritten in the final pass
mkdir -p gpurun_out
export GB200_SANITIZE=1
timeout 1700 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 --log-file gpurun_out/r4_racecheck.log \
  python -m pytest tests/test_qoix_gpu.py tests/test_qoix_encode_gpu.py tests/test_jpeg_gpu.py -m gpu -x -q \
  -k "(plane10 or depth_maps or every_opcode or restart or subsampling or grey) and not config" > gpurun_out/r4_racecheck_pytest.txt 2>&1
echo "exit $?" >> gpurun_out/r4_racecheck_pytest.txt
tail -4 gpurun_out/r4_racecheck_pytest.txt
grep -c "hazard" gpurun_out/r4_racecheck.log
grep -m 8 -B2 -A10 "hazard" gpurun_out/r4_racecheck.log | head -80
tail -3 gpurun_out/r4_racecheck.log
