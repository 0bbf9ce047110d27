#!/bin/bash
# SASS evidence for profiles/sass/: per-kernel opcode histograms of the hot kernels, the full listing of the convert
# kernel and one step of the unfilter wavefront (built objects under gamut_b200/build; no GPU needed)
set -e
cd "$(dirname "$0")/.."
O=profiles/sass; mkdir -p $O
H=$O/r2_opcode_histograms.txt
echo "# cuobjdump -sass opcode histograms (sm_100a), $(date -u +%F)" > $H
for pair in "jpeg.o:jpeg_unstuff_count_kernel" "jpeg.o:jpeg_unstuff_write_kernel" "jpeg.o:jpeg_sync_kernel" "jpeg.o:jpeg_write_kernel" "jpeg.o:jpeg_idct_colour_kernel" "jpeg.o:jpeg_prog_kernel" \
            "png_kernels.o:infp_find_kernel" "png_kernels.o:infp_count_kernel" "png_kernels.o:infp_write_kernel" "png_kernels.o:infp_resolve_kernel" \
            "png_kernels.o:unfilter_kernelILb0" "png_kernels.o:unfilter_rowpar_kernel" "qoix.o:lz4_sync_kernel" "qoix.o:lz4_pwrite_kernel" "qoix.o:lz4_resolve_kernel" \
            "qoix.o:p10_sync_kernel" "qoix.o:p10_write_kernel" "qoix.o:p10_recon_kernelILi2" "qoix.o:qoi_kernel" "qoix_encode.o:qe_tile_kernelILb0" "qoix_encode.o:qe_tile_kernelILb1" "convert.o:convert_directILi2ELi14" "convert.o:convert_directILi14ELi2"; do
  o=${pair%%:*}; k=${pair##*:}
  fn=$(cuobjdump -sass gamut_b200/build/$o | grep "Function :" | grep "$k" | head -1 | awk '{print $3}')
  [ -z "$fn" ] && continue
  cuobjdump -sass -fun "$fn" gamut_b200/build/$o | grep -o "^ *\/\*[0-9a-f]*\*\/ *[^;]*;" | sed 's/^ *\/\*[0-9a-f]*\*\/ *//' > /tmp/k.sass
  echo >> $H; echo "$k: $(wc -l < /tmp/k.sass) instructions" >> $H
  sed 's/^@!*U*P[0-9T] //' /tmp/k.sass | awk '{print $1}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -16 | awk '{printf "  %s %s;", $1, $2} END {print ""}' >> $H
  case $k in
    convert_directILi2ELi14) cp /tmp/k.sass $O/r2_convert_direct_rgba8_to_rgbaf32.sass ;;
    p10_recon_kernelILi2) L=$(grep -n "VIMNMX" /tmp/k.sass | awk -F: 'NR==1{print $1}'); sed -n "$((L-6)),$((L+60))p" /tmp/k.sass > $O/r2_p10_recon_pixel_body.sass ;;
    unfilter_kernelILb0) L=$(grep -n "SHFL.UP PT" /tmp/k.sass | awk -F: 'NR==30{print $1}'); sed -n "${L},$((L+58))p" /tmp/k.sass > $O/r2_unfilter_wavefront_one_step.sass ;;
  esac
done
grep -c . $H
