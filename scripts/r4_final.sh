#!/bin/bash
# final pass of round 2: state pass, DRAM traffic per kernel group, full ncu captures of the rewritten kernels
# (summarised on the box: the .ncu-rep files are too large to bring back)
mkdir -p gpurun_out
bash scripts/r4_state.sh r4f
timeout 1500 python scripts/capture_traffic.py > gpurun_out/r4f_traffic.log 2>&1; tail -3 gpurun_out/r4f_traffic.log
rm -f gpurun_out/traffic_*.csv
mkdir -p /tmp/cub && (cd /tmp/cub && cuobjdump -xelf all $GRAFT_REPO_ROOT/gamut_b200/libgamut_b200.so > /dev/null 2>&1)
prof() {  # $1 tag, $2 kernel regex, $3 count, rest: command
  local tag=$1 re=$2 cnt=$3; shift 3
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:"$re" -c $cnt -f -o /tmp/prof_$tag "$@" > gpurun_out/r4f_prof_$tag.log 2>&1
  python profiles/summarize.py full /tmp/prof_$tag.ncu-rep gpurun_out/r4f_ncu_full_$tag.txt
  for k in $(echo "$re" | tr '|' ' '); do
    cub=$(grep -l "$k" /tmp/cub/*.cubin 2>/dev/null | head -1)
    [ -z "$cub" ] && continue
    nvdisasm -g -c $cub > /tmp/k.sass 2>/dev/null
    ncu -i /tmp/prof_$tag.ncu-rep --page source --csv --kernel-name regex:$k > /tmp/src.csv 2>/dev/null
    echo "== $k" >> gpurun_out/r4f_hotlines_$tag.txt
    python profiles/hotlines.py /tmp/src.csv /tmp/k.sass $k 14 >> gpurun_out/r4f_hotlines_$tag.txt 2>&1
  done
  rm -f /tmp/prof_$tag.ncu-rep
}
B="python bench.py --only --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0"
prof jpeg 'jpeg_unstuff_write_kernel|jpeg_sync_kernel|jpeg_write_kernel|jpeg_idct_colour_kernel' 8 $B --workload jpeg --batch 512 --sub-batch 512
prof qoix 'lz4_sync_kernel|lz4_pwrite_kernel|lz4_resolve_kernel|p10_sync_kernel|p10_write_kernel|p10_recon_kernel' 14 $B --workload qoix --batch 148
prof png 'infp_count_kernel|infp_write_kernel' 4 $B --workload png --batch 512
prof encode 'qe_tile_kernel' 4 python scripts/qoix_encode_bench.py 64
du -sh gpurun_out
