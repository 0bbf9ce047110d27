#!/bin/bash
# one ncu --set full capture of the JPEG kernels on a small batch (64 images), raw CSV pages back in gpurun_out/
T=${1:-r2c}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'jpeg_(sync|write|idct_colour)_kernel' -s 9 -c 3 -o gpurun_out/${T}_jpeg_prof -f \
  python bench.py --workload jpeg --only --batch 64 --sub-batch 64 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${T}_ncu_full.log 2>&1
tail -3 gpurun_out/${T}_ncu_full.log
ls -la gpurun_out/${T}_jpeg_prof.ncu-rep
