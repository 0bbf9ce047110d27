// qoi_encode.cuh -- the kernels of qoi_encode.cu (see the comment at the top of that file for the formulation).
// Kept apart from the host code so that tests/test_qoi_encode_emulated.py can compile exactly this text for the host
// under a thread-per-CUDA-thread emulation (tests/cuda_emu.h) and compare it with the oracle without a GPU.
#pragma once
#include <stdint.h>
#include <stddef.h>

namespace {

constexpr int QN_TILE = 1024, QN_THREADS = 256, QN_PER = QN_TILE / QN_THREADS;
constexpr int QOI_HEADER_SIZE = 14, QOI_PADDING = 8;
constexpr uint32_t QOI_PIXELS_MAX = 400000000u;

struct QnImage {
    const uint8_t* pixels; int pitch;         // rgb8 / rgba8 rows; the pitch may be negative (pixels = first scanline)
    uint32_t w, h, np; int channels; int word_loads;
    uint32_t tile_base, ntiles;
    uint8_t* out;                             // 14-byte header + codes + 8-byte padding
    uint8_t header[QOI_HEADER_SIZE];
};
struct QnTile { int last_ne; int carry_ne; uint32_t bytes; uint32_t byte_base; uint32_t okmask[2]; };

// pixel i as r | g << 8 | b << 16 | a << 24 (a = 255 for rgb8, as px_prev's alpha stays 255 in the reference, :327-331)
__device__ __forceinline__ uint32_t qn_load(const QnImage& im, uint32_t i)
{
    const uint32_t y = i / im.w, x = i - y * im.w;
    const uint8_t* p = im.pixels + (ptrdiff_t)im.pitch * (ptrdiff_t)y + (size_t)x * im.channels;
    if (im.channels == 4) {
        if (im.word_loads) return __ldg((const uint32_t*)p);
        return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
    }
    return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | 0xff000000u;
}

// s_v[k] = pixel tile0 - 1 + k for k = 0 .. QN_TILE + 1; the pixel before the first is the initial px_prev (0,0,0,255)
__device__ __forceinline__ void qn_stage(const QnImage& im, uint32_t tile0, uint32_t* s_v)
{
    for (uint32_t k = threadIdx.x; k < QN_TILE + 2; k += QN_THREADS) {
        const long long i = (long long)tile0 - 1 + k;
        uint32_t v = 0xff000000u;
        if (i >= 0 && i < (long long)im.np) v = qn_load(im, (uint32_t)i);
        s_v[k] = v;
    }
}

__device__ __forceinline__ uint32_t qn_hash(uint32_t v)         // QOI_COLOR_HASH (:229)
{
    return ((v & 255u) * 3u + ((v >> 8) & 255u) * 5u + ((v >> 16) & 255u) * 7u + (v >> 24) * 11u) & 63u;
}

// inclusive prefix over the CTA (max or sum) of one value per thread; returns the exclusive value, *total = all
template <bool MAX>
__device__ __forceinline__ int qn_cta_scan(int v, int identity, int* s_warp, int* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int n = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc = MAX ? max(inc, n) : inc + n; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int off = identity, tot = identity;
#pragma unroll
    for (int w = 0; w < QN_THREADS / 32; ++w) { const int c = s_warp[w]; if (w < warp) off = MAX ? max(off, c) : off + c; tot = MAX ? max(tot, c) : tot + c; }
    int ex = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) ex = identity;
    __syncthreads();
    if (total) *total = tot;
    return MAX ? max(off, ex) : off + ex;
}

// ---- N1: per tile, the last pixel that differs from its predecessor and the last non-run pixel of every bucket --------
__global__ void __launch_bounds__(QN_THREADS)
qn_tile_state_kernel(const QnImage* __restrict__ imgs, QnTile* __restrict__ tiles, uint32_t* __restrict__ tile_val)
{
    __shared__ uint32_t s_v[QN_TILE + 2];
    __shared__ int s_last[64];
    __shared__ int s_warp[QN_THREADS / 32];
    const QnImage& im = imgs[blockIdx.y];
    if (blockIdx.x >= im.ntiles) return;
    const uint32_t tile_index = im.tile_base + blockIdx.x;
    const uint32_t tile0 = blockIdx.x * QN_TILE;
    qn_stage(im, tile0, s_v);
    if (threadIdx.x < 64) s_last[threadIdx.x] = -1;
    __syncthreads();
    int last = -1;
#pragma unroll
    for (int q = 0; q < QN_PER; ++q) {
        const uint32_t k = threadIdx.x * QN_PER + q;           // pixel tile0 + k = s_v[k + 1]
        if (tile0 + k < im.np) {
            const uint32_t v = s_v[k + 1];
            if (v != s_v[k]) { last = (int)(tile0 + k); atomicMax(&s_last[qn_hash(v)], (int)k); }
        }
    }
    int tot;
    qn_cta_scan<true>(last, -1, s_warp, &tot);                  // its barriers also publish s_last
    if (threadIdx.x == 0) tiles[tile_index].last_ne = tot;
    if (threadIdx.x < 64) {
        const int k = s_last[threadIdx.x];
        const uint32_t m = __ballot_sync(0xffffffffu, k >= 0);
        tile_val[(size_t)tile_index * 64 + threadIdx.x] = k >= 0 ? s_v[k + 1] : 0u;
        if ((threadIdx.x & 31) == 0) tiles[tile_index].okmask[threadIdx.x >> 5] = m;
    }
}

// ---- N2 / N4: per image, exclusive prefix over its tiles (one CTA per image) -----------------------------------------
// phase 0: prefix maximum of last_ne -> carry_ne; last writer per bucket carried forward: tile_val[t][b] becomes the
//          content of index[b] at the start of tile t (0 before the first writer, the reference's memset, :322).
// phase 1: prefix sum of bytes -> byte_base, then header, padding (:421-424) and the stream length.
__global__ void __launch_bounds__(QN_THREADS)
qn_scan_kernel(const QnImage* __restrict__ imgs, QnTile* __restrict__ tiles, uint32_t* __restrict__ tile_val, int phase,
               int* __restrict__ out_len)
{
    __shared__ int s_warp[QN_THREADS / 32];
    const QnImage& im = imgs[blockIdx.x];
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
        if (threadIdx.x < 64) {
            const uint32_t b = threadIdx.x;
            uint32_t* V = tile_val + (size_t)im.tile_base * 64 + b;
            uint32_t state = 0;
            constexpr int U = 8;
            for (uint32_t t0 = 0; t0 < im.ntiles; t0 += U) {
                uint32_t val[U], ok[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t t = t0 + u;
                    val[u] = 0; ok[u] = 0;
                    if (t < im.ntiles) { val[u] = V[(size_t)t * 64]; ok[u] = (T[t].okmask[b >> 5] >> (b & 31u)) & 1u; }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t t = t0 + u;
                    if (t < im.ntiles) { V[(size_t)t * 64] = state; if (ok[u]) state = val[u]; }
                }
            }
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
    if (threadIdx.x < QOI_HEADER_SIZE) im.out[threadIdx.x] = im.header[threadIdx.x];
    if (threadIdx.x < QOI_PADDING) im.out[QOI_HEADER_SIZE + carry + threadIdx.x] = threadIdx.x == QOI_PADDING - 1 ? 1 : 0;
    if (threadIdx.x == 0) out_len[blockIdx.x] = QOI_HEADER_SIZE + (int)carry + QOI_PADDING;
}

// ---- N3 / N5: codes of a tile. EMIT = false: bytes of the tile. EMIT = true: the bytes at their place ----------------
template <bool EMIT>
__global__ void __launch_bounds__(QN_THREADS)
qn_tile_kernel(const QnImage* __restrict__ imgs, QnTile* __restrict__ tiles, const uint32_t* __restrict__ tile_val)
{
    __shared__ uint32_t s_v[QN_TILE + 2];
    __shared__ uint32_t s_occ[64][QN_TILE / 32];               // bit k of row b: pixel k of the tile is a non-run pixel of bucket b
    __shared__ int s_warp[QN_THREADS / 32];
    __shared__ __align__(4) uint8_t s_out[EMIT ? (QN_TILE * 5 + 8) : 4];
    const QnImage& im = imgs[blockIdx.y];
    if (blockIdx.x >= im.ntiles) return;
    const uint32_t tile_index = im.tile_base + blockIdx.x;
    const QnTile tile = tiles[tile_index];
    const uint32_t tile0 = blockIdx.x * QN_TILE;
    qn_stage(im, tile0, s_v);
    for (uint32_t k = threadIdx.x; k < 64 * (QN_TILE / 32); k += QN_THREADS) (&s_occ[0][0])[k] = 0;
    __syncthreads();
    int my_last = -1;
#pragma unroll
    for (int q = 0; q < QN_PER; ++q) {
        const uint32_t k = threadIdx.x * QN_PER + q;
        if (tile0 + k < im.np) {
            const uint32_t v = s_v[k + 1];
            if (v != s_v[k]) { my_last = (int)(tile0 + k); atomicOr(&s_occ[qn_hash(v)][k >> 5], 1u << (k & 31u)); }
        }
    }
    int last_ne = max(tile.carry_ne, qn_cta_scan<true>(my_last, -1, s_warp, nullptr));     // barriers publish s_occ
    unsigned long long codes[QN_PER]; int nb[QN_PER]; int mybytes = 0;
#pragma unroll
    for (int q = 0; q < QN_PER; ++q) {
        const uint32_t k = threadIdx.x * QN_PER + q, i = tile0 + k;
        codes[q] = 0; nb[q] = 0;
        if (i >= im.np) continue;
        const uint32_t v = s_v[k + 1], pv = s_v[k];
        if (v == pv) {
            // QOI_OP_RUN (:349-356): one byte at the last pixel of a run of at most 62
            const uint32_t r = (i - (uint32_t)(last_ne + 1)) % 62u;
            const bool end = r == 61u || i + 1 == im.np || s_v[k + 2] != v;
            if (end) { codes[q] = 0xc0u | r; nb[q] = 1; }
        } else {
            last_ne = (int)i;
            const uint32_t hsh = qn_hash(v);
            // latest earlier non-run pixel of this bucket: in the tile, else what the tiles before left in the slot
            int wd = (int)(k >> 5);
            uint32_t m = s_occ[hsh][wd] & ((1u << (k & 31u)) - 1u);
            while (!m && wd > 0) m = s_occ[hsh][--wd];
            const uint32_t slot = m ? s_v[(wd << 5) + (31 - __clz(m)) + 1] : __ldg(tile_val + (size_t)tile_index * 64 + hsh);
            if (slot == v) { codes[q] = hsh; nb[q] = 1; }                                        // QOI_OP_INDEX (:372)
            else if ((v >> 24) != (pv >> 24)) { codes[q] = 0xffull | (unsigned long long)v << 8; nb[q] = 5; }   // QOI_OP_RGBA (:411)
            else {
                const int vr = (int)(signed char)((v & 255u) - (pv & 255u));
                const int vg = (int)(signed char)(((v >> 8) & 255u) - ((pv >> 8) & 255u));
                const int vb = (int)(signed char)(((v >> 16) & 255u) - ((pv >> 16) & 255u));
                const int vg_r = (int)(signed char)(vr - vg), vg_b = (int)(signed char)(vb - vg);
                if (vr > -3 && vr < 2 && vg > -3 && vg < 2 && vb > -3 && vb < 2) {                 // QOI_OP_DIFF (:388)
                    codes[q] = 0x40u | (uint32_t)(vr + 2) << 4 | (uint32_t)(vg + 2) << 2 | (uint32_t)(vb + 2); nb[q] = 1;
                } else if (vg_r > -9 && vg_r < 8 && vg > -33 && vg < 32 && vg_b > -9 && vg_b < 8) { // QOI_OP_LUMA (:396)
                    codes[q] = (0x80u | (uint32_t)(vg + 32)) | ((uint32_t)(vg_r + 8) << 4 | (uint32_t)(vg_b + 8)) << 8; nb[q] = 2;
                } else { codes[q] = 0xfeull | (unsigned long long)(v & 0xffffffu) << 8; nb[q] = 4; }   // QOI_OP_RGB (:404)
            }
        }
        mybytes += nb[q];
    }
    int total;
    const int ex = qn_cta_scan<false>(mybytes, 0, s_warp, &total);
    if (!EMIT) { if (threadIdx.x == 0) tiles[tile_index].bytes = (uint32_t)total; return; }
    // the tile's bytes are put together in shared memory at the alignment they have in memory (s_out[0] = the first
    // byte of the aligned 32-bit word the tile starts in); whole words are stored as words, the shared ends as bytes
    const uint32_t g0 = QOI_HEADER_SIZE + tile.byte_base;
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
inline bool qn_valid(uint32_t width, uint32_t height, int channels, int colorspace)      // qoi_encode's own checks (:303-311)
{
    return width && height && channels >= 3 && channels <= 4 && colorspace >= 0 && colorspace <= 1 && height < QOI_PIXELS_MAX / width;
}
// Fills the table entry of one image; false = the encoder refuses it. total_tiles is advanced by the image's tiles.
inline bool qn_setup(QnImage& Q, const uint8_t* pixels, uint32_t width, uint32_t height, int channels, int colorspace, int pitch,
                     uint8_t* out, uint32_t& total_tiles)
{
    if (!qn_valid(width, height, channels, colorspace) || !pixels || !out || ((uintptr_t)out & 15)) return false;
    const long long row = (long long)width * channels, ap = pitch < 0 ? -(long long)pitch : pitch;
    if (ap < row && height > 1) return false;                       // rows would overlap
    Q = QnImage();
    Q.pixels = pixels; Q.pitch = pitch; Q.w = width; Q.h = height; Q.np = width * height; Q.channels = channels;
    Q.word_loads = channels == 4 && ((uintptr_t)pixels & 3) == 0 && (pitch & 3) == 0;
    Q.tile_base = total_tiles; Q.ntiles = (Q.np + QN_TILE - 1) / QN_TILE; total_tiles += Q.ntiles;
    Q.out = out;
    uint8_t* h = Q.header;
    const uint32_t words[3] = {0x716F6966u, width, height};         // "qoif" (:232), big-endian (:236-241)
    for (int k = 0; k < 3; ++k) { h[4 * k] = (uint8_t)(words[k] >> 24); h[4 * k + 1] = (uint8_t)(words[k] >> 16); h[4 * k + 2] = (uint8_t)(words[k] >> 8); h[4 * k + 3] = (uint8_t)words[k]; }
    h[12] = (uint8_t)channels; h[13] = (uint8_t)colorspace;
    return true;
}

}  // namespace
