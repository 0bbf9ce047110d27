#!/bin/bash
# full ncu capture of the JPEG kernels (one launch each) on a 512-image sub-batch
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'jpeg_sync_kernel|jpeg_write_kernel|jpeg_idct_colour_kernel|jpeg_unstuff_kernel' -c 4 -f -o gpurun_out/prof_jpeg_r3 python bench.py --workload jpeg --only --batch 512 --sub-batch 512 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/prof_jpeg_r3.log 2>&1
tail -2 gpurun_out/prof_jpeg_r3.log | cut -c1-200
ls -la gpurun_out/prof_jpeg_r3.ncu-rep
