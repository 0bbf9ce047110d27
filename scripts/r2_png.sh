#!/bin/bash
# PNG iteration pass: inflate/png parity, launch list, bench line
T=${1:-r2r}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_inflate_gpu.py tests/test_png_gpu.py tests/test_image_gpu.py tests/test_batch_host_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${T}_pytest_png.txt
cat gpurun_out/${T}_pytest_png.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches_png.csv python bench.py --workload png --only --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${T}_ncu_png.log 2>&1
python scripts/launch_summary.py gpurun_out/${T}_launches_png.csv | grep -v "at::"
timeout 600 python bench.py --workload png --only --no-cpu-baseline > gpurun_out/${T}_bench_png.json 2> gpurun_out/${T}_bench_png.err
cut -c1-300 gpurun_out/${T}_bench_png.json; tail -3 gpurun_out/${T}_bench_png.err
timeout 120 python scripts/hw_decompress_bar.py > gpurun_out/hw_bar.json 2> gpurun_out/hw_bar.err; cat gpurun_out/hw_bar.json; tail -3 gpurun_out/hw_bar.err
