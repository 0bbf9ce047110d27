// tga_encode.cuh -- kernels of the TGA encoder (host code in tga.cu). Compiled for the GPU by tga.cu and for the host,
// under the thread-per-CUDA-thread emulation, by tests/emu_tga.cpp.
//
// Reference: saveTGA (plugins/tga.d:123-149) -> TGAEncoder.encodeScanline (codecs/tga.d:144-292), run-length coding on.
// Every scanline is coded on its own, so rows are the parallel axis: one warp per row.
//   * similarMask[x] = pixel x equals pixel x - 1 (:186-203): one ballot per 32 pixels, kept as a bit mask in shared memory;
//   * the backward pass (:207-243) gives every position the packet that would start there: a run when the next pixel is
//     similar (0x80 | number of similar pixels that follow, at most 127), else a raw packet (the number of following
//     pixels that differ from their predecessors, at most 127). Its float comparison only ever compares a finite value
//     with an infinite one (or two infinities at the last pixel, where "<=" picks raw), so the choice is "run iff the
//     next pixel is similar". Both counts are bit scans over the mask;
//   * the forward pass (:252-271) follows packet to packet: the warp walks that chain together and its lanes copy the
//     pixels of the packet.
// Two passes of the same kernel: bytes per row, then (after a prefix sum over the rows of an image) the bytes themselves.
#pragma once
#include <stdint.h>
#include <stddef.h>

namespace {

constexpr int TE_MAX_W = 65535;                       // TGAEncoder.initialize refuses more (:78-81)
constexpr int TE_MASK_WORDS = (TE_MAX_W + 32) / 32;

struct TeImage {
    const uint8_t* pixels; int pitch;                 // first scanline, signed pitch; l8 / la8 / rgb8 / rgba8
    int w, h, src_channels, channels;                 // channels of the file: 3 (l8, rgb8) or 4 (la8, rgba8)
    uint32_t row_base;                                // first entry of this image in the row table
    uint8_t* out;                                     // 18-byte header + rows, bottom row first
};

// pixel x of scanline `scan` as the file stores it: b | g << 8 | r << 16 | a << 24 (scanline_convert_*_to_rgb8 / rgba8
// then the R / B swap, :152-178; a = 255 for 24-bit files, as in the comparison of :196)
__device__ __forceinline__ uint32_t te_px(const TeImage& im, const uint8_t* scan, int x)
{
    const uint8_t* p = scan + (size_t)x * im.src_channels;
    if (im.src_channels == 1) return (uint32_t)p[0] * 0x010101u | 0xff000000u;
    if (im.src_channels == 2) return (uint32_t)p[0] * 0x010101u | (uint32_t)p[1] << 24;
    if (im.src_channels == 3) return (uint32_t)p[2] | (uint32_t)p[1] << 8 | (uint32_t)p[0] << 16 | 0xff000000u;
    return (uint32_t)p[2] | (uint32_t)p[1] << 8 | (uint32_t)p[0] << 16 | (uint32_t)p[3] << 24;
}

// number of consecutive mask bits equal to `val` from position p on, at most 127 and not past w
__device__ __forceinline__ int te_run(const uint32_t* mask, int p, int w, uint32_t val)
{
    int c = 0;
    while (c < 127 && p + c < w) {
        const int q = p + c;
        uint32_t word = mask[q >> 5];
        if (!val) word = ~word;
        word >>= (q & 31);
        const int avail = 32 - (q & 31);
        const uint32_t inv = ~word;
        int t = inv ? __ffs((int)inv) - 1 : 32;       // trailing bits that match
        if (t > avail) t = avail;
        c += t;
        if (t < avail) break;
    }
    if (c > 127) c = 127;
    if (c > w - p) c = w - p;
    return c;
}

// EMIT = false: row_bytes[row] = bytes of the row. EMIT = true: the bytes, at out + 18 + row_off[row].
// Row r of the table is file row r of its image = scanline h - 1 - r (plugins/tga.d:141-145).
template <bool EMIT>
__global__ void __launch_bounds__(32)
te_row_kernel(const TeImage* __restrict__ imgs, uint32_t* __restrict__ row_bytes, const uint32_t* __restrict__ row_off)
{
    __shared__ uint32_t s_mask[TE_MASK_WORDS];
    const TeImage& im = imgs[blockIdx.y];
    if ((int)blockIdx.x >= im.h) return;
    const int lane = threadIdx.x, w = im.w, ch = im.channels;
    const uint8_t* scan = im.pixels + (ptrdiff_t)im.pitch * (ptrdiff_t)(im.h - 1 - (int)blockIdx.x);
    for (int x0 = 0; x0 < w; x0 += 32) {
        const int x = x0 + lane;
        const bool similar = x < w && x > 0 && te_px(im, scan, x) == te_px(im, scan, x - 1);
        const uint32_t m = __ballot_sync(0xffffffffu, similar);
        if (lane == 0) s_mask[x0 >> 5] = m;
    }
    __syncwarp();
    uint8_t* out = EMIT ? im.out + 18 + row_off[im.row_base + blockIdx.x] : nullptr;
    uint32_t bytes = 0;
    int x = 0;
    while (x < w) {                                   // every lane walks the same chain
        const bool run = x + 1 < w && ((s_mask[(x + 1) >> 5] >> ((x + 1) & 31)) & 1u);
        const int follow = te_run(s_mask, x + 1, w, run ? 1u : 0u);
        const int n = follow + 1;
        if (EMIT) {
            uint8_t* o = out + bytes;
            if (lane == 0) o[0] = (uint8_t)(run ? 0x80 | follow : follow);
            for (int k = lane; k < (run ? 1 : n); k += 32) {
                const uint32_t v = te_px(im, scan, x + k);
                uint8_t* d = o + 1 + k * ch;
                d[0] = (uint8_t)v; d[1] = (uint8_t)(v >> 8); d[2] = (uint8_t)(v >> 16);
                if (ch == 4) d[3] = (uint8_t)(v >> 24);
            }
        }
        bytes += 1u + (uint32_t)(run ? ch : n * ch);
        x += n;
    }
    if (!EMIT && lane == 0) row_bytes[im.row_base + blockIdx.x] = bytes;
}

// per image: exclusive prefix sum of its rows' bytes (one CTA of 256 threads per image), the header (:120-131) and the
// file length
__global__ void __launch_bounds__(256)
te_scan_kernel(const TeImage* __restrict__ imgs, const uint32_t* __restrict__ row_bytes, uint32_t* __restrict__ row_off,
               int* __restrict__ out_len)
{
    __shared__ uint32_t s_warp[8];
    const TeImage& im = imgs[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t carry = 0;
    for (int r0 = 0; r0 < im.h; r0 += 256) {
        const int r = r0 + (int)threadIdx.x;
        const uint32_t v = r < im.h ? row_bytes[im.row_base + r] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t off = 0, tot = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { const uint32_t c = s_warp[k]; if (k < warp) off += c; tot += c; }
        if (r < im.h) row_off[im.row_base + r] = carry + off + inc - v;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x < 18) {
        const int k = threadIdx.x;
        uint8_t b = 0;
        if (k == 2) b = 10;
        else if (k == 12) b = (uint8_t)(im.w & 0xff);
        else if (k == 13) b = (uint8_t)((im.w & 0xff00) >> 8);
        else if (k == 14) b = (uint8_t)(im.h & 0xff);
        else if (k == 15) b = (uint8_t)((im.h & 0xff00) >> 8);
        else if (k == 16) b = (uint8_t)(im.channels * 8);
        im.out[k] = b;
    }
    if (threadIdx.x == 0) out_len[blockIdx.x] = 18 + (int)carry;
}

// ---- host side of the image table (shared with the emulation harness) ----------------------------------------------
// type = PixelType value (types.d): l8 = 0, la8 = 3, rgb8 = 9, rgba8 = 12 (TGAEncoder.initialize, :84-111)
inline int te_src_channels(int type) { return type == 0 ? 1 : type == 3 ? 2 : type == 9 ? 3 : type == 12 ? 4 : 0; }
inline size_t te_bound(int type, int width, int height)
{
    const int sc = te_src_channels(type);
    if (!sc || width < 0 || height < 0) return 0;
    const size_t ch = (sc & 1) ? 3 : 4;
    return 18 + (size_t)height * ((size_t)width * ch + (size_t)width / 2 + 2) + 16;
}
inline bool te_setup(TeImage& T, const uint8_t* pixels, int type, int width, int height, int pitch, uint8_t* out, uint32_t& total_rows)
{
    const int sc = te_src_channels(type);
    if (!sc || !pixels || !out || width < 1 || height < 1 || width > TE_MAX_W || height > 65535) return false;
    if (te_bound(type, width, height) > 0x7fffffffull) return false;          // row offsets and the length are 32-bit
    const long long ap = pitch < 0 ? -(long long)pitch : pitch;
    if (ap < (long long)width * sc && height > 1) return false;
    T = TeImage();
    T.pixels = pixels; T.pitch = pitch; T.w = width; T.h = height; T.src_channels = sc; T.channels = (sc & 1) ? 3 : 4;
    T.row_base = total_rows; total_rows += (uint32_t)height;
    T.out = out;
    return true;
}

}  // namespace
