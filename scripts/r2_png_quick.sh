#!/bin/bash
# quick PNG pass: inflate parity + launch list
T=${1:-r2s}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_inflate_gpu.py tests/test_png_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches_png.csv python bench.py --workload png --only --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${T}_ncu_png.log 2>&1
python scripts/launch_summary.py gpurun_out/${T}_launches_png.csv | grep -v "at::"
