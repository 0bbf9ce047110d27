// qoi2avg_encode.cuh -- kernels of the QOI2AVG encoder (qoix_encode, codecs/qoi2avg.d:376-617: the codec saveQOIX picks
// for rgb8 / rgba8 images). Host code in qoi2avg_encode.cu; compiled for the host under the thread-per-CUDA-thread
// emulation by tests/emu_qoi2avg_encode.cpp.
//
// Everything in this encoder is a function of a few input pixels -- the run position (prefix maximum, runs cut every
// 1024), the alpha difference, the LOCO-I prediction from the ORIGINAL left / above / above-left pixels, the choice
// between LUMA / GRAY / LUMA2 / LUMA3 / RGB -- except the colour index. Unlike plain QOI's, this index is a 64-entry FIFO
// that only MISSES enter (:496-507): whether a pixel hits depends on how many earlier pixels missed (an entry dies
// after 64 further misses) and the emitted byte carries the FIFO slot, i.e. the miss count modulo 64. That chain is
// serial, but it is small: Q1 below walks it with one thread per image and leaves one byte per pixel (0 = miss,
// 0x80 | slot = the INDEX opcode itself); the other four kernels are the tile-parallel shape of the QOI encoder
// (qoi_encode.cuh, whose helpers are reused): last differing pixel per tile, prefix maximum, bytes per tile, prefix sum
// (+ header, padding, length), emit.
#pragma once
#include "qoi_encode.cuh"

namespace {

constexpr int Q2_HEADER_SIZE = 25, Q2_PADDING = 4;

struct Q2Image {
    QnImage base;                 // pixels, pitch, w, h, np, channels, word_loads, tile_base, ntiles, out (header[] unused)
    uint8_t* ix;                  // np bytes: 0 = not an index hit, 0x80 | slot = QOI_OP_INDEX byte (written by Q1)
    uint8_t header[Q2_HEADER_SIZE];
};

__device__ __forceinline__ uint32_t q2_hash(uint32_t v) { return ((v * 2654435769u) >> 22) & 1023u; }      // QOI_COLOR_HASH (:312-315)

// ---- Q1: the colour index, one warp per image. The lanes stage 1024 pixels at a time in shared memory and store the
// result bytes; lane 0 walks the pixels (:472-507: run pixels do not touch the index; a pixel found in
// index[index_lookup[hash]] is a hit; any other pixel takes the next FIFO slot).
__global__ void __launch_bounds__(32)
q2_index_kernel(const Q2Image* __restrict__ imgs)
{
    __shared__ uint32_t s_v[QN_TILE + 1];
    __shared__ uint8_t s_ix[QN_TILE];
    __shared__ uint32_t s_index[64];
    __shared__ uint8_t s_lookup[1024];
    const Q2Image& im = imgs[blockIdx.x];
    const uint32_t lane = threadIdx.x, np = im.base.np;
    for (uint32_t k = lane; k < 64; k += 32) s_index[k] = 0;
    for (uint32_t k = lane; k < 1024; k += 32) s_lookup[k] = 0;
    uint32_t index_pos = 0;
    for (uint32_t t0 = 0; t0 < np; t0 += QN_TILE) {
        const uint32_t n = min((uint32_t)QN_TILE, np - t0);
        __syncwarp();
        for (uint32_t k = lane; k < n + 1; k += 32)               // s_v[k] = pixel t0 - 1 + k; before the first: (0,0,0,255) (:438-441)
            s_v[k] = (t0 + k == 0) ? 0xff000000u : qn_load(im.base, t0 + k - 1);
        __syncwarp();
        if (lane == 0) {
            for (uint32_t k = 0; k < n; ++k) {
                const uint32_t v = s_v[k + 1];
                uint8_t r = 0;
                if (v != s_v[k]) {
                    const uint32_t h = q2_hash(v);
                    const uint32_t slot = s_lookup[h];
                    if (s_index[slot] == v) r = (uint8_t)(0x80u | slot);
                    else { s_lookup[h] = (uint8_t)index_pos; s_index[index_pos] = v; index_pos = (index_pos + 1u) & 63u; }
                }
                s_ix[k] = r;
            }
        }
        __syncwarp();
        for (uint32_t k = lane; k < n; k += 32) im.ix[t0 + k] = s_ix[k];
    }
}

// ---- Q2: per tile, the last pixel that differs from its predecessor -------------------------------------------------
__global__ void __launch_bounds__(QN_THREADS)
q2_tile_ne_kernel(const Q2Image* __restrict__ imgs, QnTile* __restrict__ tiles)
{
    __shared__ uint32_t s_v[QN_TILE + 2];
    __shared__ int s_warp[QN_THREADS / 32];
    const QnImage& im = imgs[blockIdx.y].base;
    if (blockIdx.x >= im.ntiles) return;
    const uint32_t tile0 = blockIdx.x * QN_TILE;
    qn_stage(im, tile0, s_v);
    __syncthreads();
    int last = -1;
#pragma unroll
    for (int q = 0; q < QN_PER; ++q) {
        const uint32_t k = threadIdx.x * QN_PER + q;
        if (tile0 + k < im.np && s_v[k + 1] != s_v[k]) last = (int)(tile0 + k);
    }
    int tot;
    qn_cta_scan<true>(last, -1, s_warp, &tot);
    if (threadIdx.x == 0) tiles[im.tile_base + blockIdx.x].last_ne = tot;
}

// ---- Q3 / Q5: per image, exclusive prefix over its tiles. phase 0: prefix maximum of last_ne -> carry_ne. phase 1:
// prefix sum of bytes -> byte_base, then header (:417-430), padding (:608-611) and the stream length.
__global__ void __launch_bounds__(QN_THREADS)
q2_scan_kernel(const Q2Image* __restrict__ imgs, QnTile* __restrict__ tiles, int phase, int* __restrict__ out_len)
{
    __shared__ int s_warp[QN_THREADS / 32];
    const Q2Image& qi = imgs[blockIdx.x];
    const QnImage& im = qi.base;
    QnTile* T = tiles + im.tile_base;
    if (phase == 0) {
        int carry = -1;
        for (uint32_t t0 = 0; t0 < im.ntiles; t0 += QN_THREADS) {
            const uint32_t t = t0 + threadIdx.x;
            const int v = t < im.ntiles ? T[t].last_ne : -1;
            int tot;
            const int ex = qn_cta_scan<true>(v, -1, s_warp, &tot);
            if (t < im.ntiles) T[t].carry_ne = max(carry, ex);
            carry = max(carry, tot);
        }
        return;
    }
    uint32_t carry = 0;
    for (uint32_t t0 = 0; t0 < im.ntiles; t0 += QN_THREADS) {
        const uint32_t t = t0 + threadIdx.x;
        const int v = t < im.ntiles ? (int)T[t].bytes : 0;
        int tot;
        const int ex = qn_cta_scan<false>(v, 0, s_warp, &tot);
        if (t < im.ntiles) T[t].byte_base = carry + (uint32_t)ex;
        carry += (uint32_t)tot;
    }
    if (threadIdx.x < Q2_HEADER_SIZE) im.out[threadIdx.x] = qi.header[threadIdx.x];
    if (threadIdx.x < Q2_PADDING) im.out[Q2_HEADER_SIZE + carry + threadIdx.x] = 0xff;
    if (threadIdx.x == 0) out_len[blockIdx.x] = Q2_HEADER_SIZE + (int)carry + Q2_PADDING;
}

// locoIntraPredictionSIMD for one channel (:863-897): A + B - C; min(A, B) where C >= max(A, B); then max(A, B) where
// C <= min(A, B) (applied second, so it wins when both hold); saturated to 0..255 by the pack
__device__ __forceinline__ int q2_loco(int a, int b, int c)
{
    const int mx = max(a, b), mn = min(a, b);
    int p = a + b - c;
    if (c >= mx) p = mn;
    if (c <= mn) p = mx;
    return min(max(p, 0), 255);
}

// ---- Q4 / Q6: codes of a tile. EMIT = false: bytes of the tile. EMIT = true: the bytes at their place ----------------
template <bool EMIT>
__global__ void __launch_bounds__(QN_THREADS)
q2_tile_kernel(const Q2Image* __restrict__ imgs, QnTile* __restrict__ tiles)
{
    __shared__ uint32_t s_v[QN_TILE + 2];
    __shared__ int s_warp[QN_THREADS / 32];
    __shared__ __align__(4) uint8_t s_out[EMIT ? (QN_TILE * 5 + 8) : 4];
    const Q2Image& qi = imgs[blockIdx.y];
    const QnImage& im = qi.base;
    if (blockIdx.x >= im.ntiles) return;
    const uint32_t tile_index = im.tile_base + blockIdx.x;
    const QnTile tile = tiles[tile_index];
    const uint32_t tile0 = blockIdx.x * QN_TILE;
    qn_stage(im, tile0, s_v);
    __syncthreads();
    int my_last = -1;
#pragma unroll
    for (int q = 0; q < QN_PER; ++q) {
        const uint32_t k = threadIdx.x * QN_PER + q;
        if (tile0 + k < im.np && s_v[k + 1] != s_v[k]) my_last = (int)(tile0 + k);
    }
    int last_ne = max(tile.carry_ne, qn_cta_scan<true>(my_last, -1, s_warp, nullptr));
    unsigned long long codes[QN_PER]; int nb[QN_PER]; int mybytes = 0;
#pragma unroll
    for (int q = 0; q < QN_PER; ++q) {
        const uint32_t k = threadIdx.x * QN_PER + q, i = tile0 + k;
        codes[q] = 0; nb[q] = 0;
        if (i >= im.np) continue;
        const uint32_t v = s_v[k + 1], pv = s_v[k];
        if (v == pv) {
            // one code at the last pixel of a run of at most 1024 (:474-482, :488-497): the short form only when a
            // different pixel ends a run of at most 8
            const uint32_t r = (i - (uint32_t)(last_ne + 1)) & 1023u;               // run - 1
            const bool forced = r == 1023u || i + 1 == im.np;
            if (forced || s_v[k + 2] != v) {
                if (!forced && r < 8u) { codes[q] = 0xf0u | r; nb[q] = 1; }
                else { codes[q] = (0xf8u | ((r >> 8) & 3u)) | (r & 0xffu) << 8; nb[q] = 2; }
            }
        } else {
            last_ne = (int)i;
            const uint32_t hit = qi.ix[i];
            if (hit) { codes[q] = hit; nb[q] = 1; }                                  // QOI_OP_INDEX (:500-502)
            else {
                const int r8 = (int)(v & 255u), g8 = (int)((v >> 8) & 255u), b8 = (int)((v >> 16) & 255u);
                const int va = (int)(signed char)((v >> 24) - (pv >> 24));
                unsigned long long c = 0; int n = 0;
                bool colour = true;
                if (va) {
                    if (va >= -4 && va <= 3) { c = 0xe8u | (uint32_t)(va + 4); n = 1; }          // QOI_OP_ADIFF (:511-512)
                    else { c = 0xfeull | (unsigned long long)v << 8; n = 5; colour = false; }    // QOI_OP_RGBA (:514-519)
                }
                if (colour) {
                    // the reference colour: the previous pixel in the first row, the pixel above in the first column,
                    // else the LOCO-I prediction of left / above / above-left (:525-545)
                    const uint32_t y = i / im.w, x = i - y * im.w;
                    int rr = (int)(pv & 255u), rg = (int)((pv >> 8) & 255u), rb = (int)((pv >> 16) & 255u);
                    if (y > 0) {
                        const uint32_t up = qn_load(im, i - im.w);
                        if (x == 0) { rr = (int)(up & 255u); rg = (int)((up >> 8) & 255u); rb = (int)((up >> 16) & 255u); }
                        else {
                            const uint32_t ul = qn_load(im, i - im.w - 1);
                            rr = q2_loco(rr, (int)(up & 255u), (int)(ul & 255u));
                            rg = q2_loco(rg, (int)((up >> 8) & 255u), (int)((ul >> 8) & 255u));
                            rb = q2_loco(rb, (int)((up >> 16) & 255u), (int)((ul >> 16) & 255u));
                        }
                    }
                    const int vg = (int)(signed char)(g8 - rg);
                    const int vg_r = (int)(signed char)(r8 - rr - vg), vg_b = (int)(signed char)(b8 - rb - vg);
                    unsigned long long cc; int cn;
                    if (vg >= -4 && vg < 0 && vg_r >= -1 && vg_r <= 2 && vg_b >= -1 && vg_b <= 2) {
                        cc = (uint32_t)(vg + 4) << 4 | (uint32_t)(vg_r + 1) << 2 | (uint32_t)(vg_b + 1); cn = 1;           // QOI_OP_LUMA (:548-554)
                    } else if (vg >= 0 && vg <= 3 && vg_r >= -2 && vg_r <= 1 && vg_b >= -2 && vg_b <= 1) {
                        cc = (uint32_t)(vg + 4) << 4 | (uint32_t)(vg_r + 2) << 2 | (uint32_t)(vg_b + 2); cn = 1;           // QOI_OP_LUMA (:555-561)
                    } else if (g8 == r8 && g8 == b8) {
                        cc = 0xfcu | (uint32_t)g8 << 8; cn = 2;                                                            // QOI_OP_GRAY (:562-568)
                    } else if (vg_r >= -8 && vg_r <= 7 && vg >= -16 && vg <= 15 && vg_b >= -8 && vg_b <= 7) {
                        cc = (0xc0u | (uint32_t)(vg + 16)) | ((uint32_t)(vg_r + 8) << 4 | (uint32_t)(vg_b + 8)) << 8; cn = 2;   // QOI_OP_LUMA2 (:569-576)
                    } else if (vg_r >= -32 && vg_r <= 31 && vg >= -64 && vg <= 63 && vg_b >= -32 && vg_b <= 31) {
                        const uint32_t dv = (uint32_t)(vg + 64) << 12 | (uint32_t)(vg_r + 32) << 6 | (uint32_t)(vg_b + 32);
                        cc = (0xe0u | ((dv >> 16) & 31u)) | ((dv >> 8) & 255u) << 8 | (dv & 255u) << 16; cn = 3;           // QOI_OP_LUMA3 (:577-586)
                    } else { cc = 0xfdull | (unsigned long long)(v & 0xffffffu) << 8; cn = 4; }                            // QOI_OP_RGB (:587-592)
                    c |= cc << (8 * n); n += cn;
                }
                codes[q] = c; nb[q] = n;
            }
        }
        mybytes += nb[q];
    }
    int total;
    const int ex = qn_cta_scan<false>(mybytes, 0, s_warp, &total);
    if (!EMIT) { if (threadIdx.x == 0) tiles[tile_index].bytes = (uint32_t)total; return; }
    // as in qn_tile_kernel: the tile's bytes at their memory alignment in shared memory, whole words stored as words
    const uint32_t g0 = Q2_HEADER_SIZE + tile.byte_base;
    const uint32_t mis = g0 & 3u;
    uint32_t p = mis + (uint32_t)ex;
#pragma unroll
    for (int q = 0; q < QN_PER; ++q) {
        unsigned long long c = codes[q];
        for (int j = 0; j < nb[q]; ++j) { s_out[p++] = (uint8_t)c; c >>= 8; }
    }
    __syncthreads();
    const uint32_t end = mis + (uint32_t)total, nwords = (end + 3u) >> 2;
    uint8_t* const base = im.out + (g0 - mis);                      // im.out is 16-byte aligned
    for (uint32_t w = threadIdx.x; w < nwords; w += QN_THREADS) {
        const uint32_t lo = max(w * 4u, mis), hi = min(w * 4u + 4u, end);
        if (hi - lo == 4u) ((uint32_t*)base)[w] = ((const uint32_t*)s_out)[w];
        else for (uint32_t b = lo; b < hi; ++b) base[b] = s_out[b];
    }
}

// ---- host side of the image table (shared with the emulation harness) ----------------------------------------------
inline bool q2_valid(uint32_t width, uint32_t height, int channels, int bitdepth, int colorspace, int compression)   // :386-398
{
    return width && height && channels >= 3 && channels <= 4 && colorspace >= 0 && colorspace <= 2 && bitdepth == 8 &&
           compression == 0 && height < QOI_PIXELS_MAX / width;
}
// Fills the table entry of one image; false = the encoder refuses it. `ix` = np bytes of scratch for this image.
inline bool q2_setup(Q2Image& Q, const uint8_t* pixels, uint32_t width, uint32_t height, int pitch, int channels, int bitdepth,
                     int colorspace, int compression, float pixelAspectRatio, float resolutionY, uint8_t* out, uint8_t* ix,
                     uint32_t& total_tiles)
{
    if (!q2_valid(width, height, channels, bitdepth, colorspace, compression) || !pixels || !out || ((uintptr_t)out & 15)) return false;
    if (pitch < (int)(width * (uint32_t)channels)) return false;   // the reference copies pitchBytes bytes of a row (:459): no flipped storage
    Q = Q2Image();
    QnImage& B = Q.base;
    B.pixels = pixels; B.pitch = pitch; B.w = width; B.h = height; B.np = width * height; B.channels = channels;
    B.word_loads = channels == 4 && ((uintptr_t)pixels & 3) == 0 && (pitch & 3) == 0;
    B.tile_base = total_tiles; B.ntiles = (B.np + QN_TILE - 1) / QN_TILE; total_tiles += B.ntiles;
    B.out = out;
    Q.ix = ix;
    uint8_t* h = Q.header;
    uint32_t fa, fr;
    __builtin_memcpy(&fa, &pixelAspectRatio, 4); __builtin_memcpy(&fr, &resolutionY, 4);
    const uint32_t w3[3] = {0x716F6978u, width, height}, w2[2] = {fa, fr};    // "qoix", big-endian
    for (int k = 0; k < 3; ++k) { h[4 * k] = (uint8_t)(w3[k] >> 24); h[4 * k + 1] = (uint8_t)(w3[k] >> 16); h[4 * k + 2] = (uint8_t)(w3[k] >> 8); h[4 * k + 3] = (uint8_t)w3[k]; }
    h[12] = 1; h[13] = (uint8_t)channels; h[14] = (uint8_t)bitdepth; h[15] = (uint8_t)colorspace; h[16] = 0;             // :420-424
    for (int k = 0; k < 2; ++k) { h[17 + 4 * k] = (uint8_t)(w2[k] >> 24); h[18 + 4 * k] = (uint8_t)(w2[k] >> 16); h[19 + 4 * k] = (uint8_t)(w2[k] >> 8); h[20 + 4 * k] = (uint8_t)w2[k]; }
    return true;
}

}  // namespace
