// png_kernels.cuh -- job descriptors shared by png.cu (host orchestration) and the PNG kernels.
#pragma once
#include <stdint.h>
#include "inflate.cuh"
#include "common.h"

namespace gb {

// One row-unfilter job = one image, or one Adam7 pass of an interlaced image.
// Replaces the row loops of stbi__create_png_image_raw (stbdec.d:1432-1547).
struct UnfilterJob {
    const uint8_t* raw;     // first filter byte of this (sub)image in the inflated stream
    uint8_t* out;           // unfiltered rows, `row_bytes` bytes each, `out_pitch` apart
    uint32_t row_bytes;     // bytes per row without the filter byte (img_width_bytes)
    uint32_t height;
    uint32_t bpp;           // filter_bytes: 1 (depth<8), else channels*bytes
    uint32_t out_pitch;
    int image;              // index into the status array
    int inflate_idx;        // index of the InflateJob that produced `raw` (-1: none)
    uint32_t need_len;      // inflated bytes the whole image needs ("not enough pixels", stbdec.d:1430)
};

// Per-image description for the elementwise "finish" kernel, which fuses every remaining step of
// the reference pipeline in source order: bit expansion + grey scaling (stbdec.d:1552-1599),
// alpha insertion (:1467-1476,1506-1545,1600-1618), 16-bit big-endian -> native (:1621-1632),
// Adam7 scatter (:1649-1676), tRNS colour key (:1682-1730), palette expansion (:1732-1765),
// stbi__convert_format[16] (:916-1200) and 16<->8 (:635-666).
struct FinishJob {
    const uint8_t* packed;        // unfiltered rows of pass p start at packed + pass_off[p]
    uint32_t pass_off[7];
    uint32_t pass_w[7], pass_h[7], pass_rb[7];
    uint8_t* out;
    uint32_t w, h;
    uint8_t depth, color, interlace, img_n;     // img_n: channels stored in the file rows
    uint8_t add_alpha;                          // out_n == img_n + 1 at unfilter time
    uint8_t has_trans, pal_n;                   // pal_n: 0 or the palette expansion width (3/4)
    uint8_t cur_n;                              // channels after palette/tRNS (img_out_n)
    uint8_t req_n;                              // final channel count
    uint8_t out16;                              // 1: final samples are 16-bit
    uint8_t tc[3];
    uint16_t tc16[3];
    uint8_t palette[1024];
};

// Scratch of one launch_inflate call; must stay alive until the stream has run the launch.
struct InflateWork { DevBuf buf; };

// h_jobs: host copy of the job table (sizes plan the scratch). Returns false when scratch allocation fails.
bool launch_inflate(InflateJob* d_jobs, const InflateJob* h_jobs, int njobs, cudaStream_t st, InflateWork& W);
void set_inflate_mode(int m);

} // namespace gb
