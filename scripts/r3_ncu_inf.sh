#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'infp_write_kernel|infp_count_kernel|infp_find_kernel' -c 3 -f -o gpurun_out/prof_inf_r3 python bench.py --workload png --only --batch 512 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/prof_inf_r3.log 2>&1
tail -1 gpurun_out/prof_inf_r3.log | cut -c1-200; ls -la gpurun_out/prof_inf_r3.ncu-rep
