// qoi10b_encode.cuh -- kernels of the QOI-10b encoder (qoi10b_encode, codecs/qoi10b.d:136-500: the codec saveQOIX picks
// for rgb16 / rgba16 images). Host code in qoi10b_encode.cu; compiled for the host under the thread-per-CUDA-thread
// emulation by tests/emu_qoi10b_encode.cpp.
//
// The reference's encoder writes stream version 1 (:168) and, although it declares a colour index (:233-235), never emits
// an index opcode: every code is a function of the pixel, its predecessor and the pixel above (prediction = rounded-up
// average of the two, :357-362). Nothing is serial but the run positions (prefix maximum, runs cut every 256) and the
// bit positions (prefix sum; codes are 2-bit aligned and up to 52 bits long: ADIFF2 + RGB). Same shape, tables and scan
// kernel as the QOI-Plane10 encoder (qoix_encode.cuh); only the per-pixel code and the pixel loads differ. Channels 3 / 4
// only: qoix_lz4_encode sends 10-bit images with 1 / 2 channels to QOI-Plane10 (plugins/qoix.d:268-278).
#pragma once
#include "qoix_encode.cuh"

namespace {

constexpr int Q10_MAX_BITS = 52;

// pixel (y, x) as r | g << 10 | b << 20 | a << 30 (10-bit values: the 16-bit samples >> 6, :288-291; a = 1023 for rgb16)
__device__ __forceinline__ unsigned long long q10_load(const QeImage& im, uint32_t y, uint32_t x)
{
    const uint16_t* p = (const uint16_t*)(im.pixels + (size_t)im.pitch * y) + (size_t)x * im.channels;
    const unsigned long long r = (uint32_t)p[0] >> 6, g = (uint32_t)p[1] >> 6, b = (uint32_t)p[2] >> 6;
    const unsigned long long a = im.channels == 4 ? (uint32_t)p[3] >> 6 : 1023u;
    return r | g << 10 | b << 20 | a << 30;
}
__device__ __forceinline__ unsigned long long q10_load_i(const QeImage& im, uint32_t i) { const uint32_t y = i / im.w; return q10_load(im, y, i - y * im.w); }
constexpr unsigned long long Q10_INITIAL = 1023ull << 30;                       // initialPredictor {0, 0, 0, 1023} (:118)

__device__ __forceinline__ bool q10_fits(uint32_t v, uint32_t k) { return v >= 1024u - k || v < k; }

// the code pixel i emits as a pixel of its own (:318-470); `above` is only read for y > 0
struct Q10Eval { bool eq; unsigned long long code; int nbits; };
__device__ __forceinline__ Q10Eval q10_eval_px(unsigned long long cur, unsigned long long prev, unsigned long long above, bool has_above)
{
    Q10Eval e;
    e.eq = cur == prev; e.code = 0; e.nbits = 0;
    if (e.eq) return e;
    const uint32_t r = (uint32_t)cur & 1023u, g = (uint32_t)(cur >> 10) & 1023u, b = (uint32_t)(cur >> 20) & 1023u, a = (uint32_t)(cur >> 30) & 1023u;
    uint32_t rr = (uint32_t)prev & 1023u, rg = (uint32_t)(prev >> 10) & 1023u, rb = (uint32_t)(prev >> 20) & 1023u;
    const uint32_t ra = (uint32_t)(prev >> 30) & 1023u;
    const uint32_t va = (a - ra) & 1023u;
    if (va) {
        if (q10_fits(va, 16)) { e.code = (0x1du << 5) | (va & 0x1fu); e.nbits = 10; }                                 // QOI_OP_ADIFF
        else if (q10_fits(va, 128)) { e.code = (0x3eu << 8) | (va & 0xffu); e.nbits = 14; }                           // QOI_OP_ADIFF2: 111110 + 8 bits
        else {                                                                                                        // QOI_OP_RGBA
            e.code = (unsigned long long)0xfeu << 40 | (unsigned long long)r << 30 | (unsigned long long)g << 20 | (unsigned long long)b << 10 | a;
            e.nbits = 48;
            return e;
        }
    }
    if (has_above) {
        rr = (rr + ((uint32_t)above & 1023u) + 1u) >> 1;
        rg = (rg + ((uint32_t)(above >> 10) & 1023u) + 1u) >> 1;
        rb = (rb + ((uint32_t)(above >> 20) & 1023u) + 1u) >> 1;
    }
    const uint32_t vg = (g - rg) & 1023u;
    const uint32_t vg_r = (r - rr - vg) & 1023u, vg_b = (b - rb - vg) & 1023u;
    unsigned long long c; int n;
    if (q10_fits(vg_r, 4) && q10_fits(vg, 8) && q10_fits(vg_b, 4)) {
        c = (unsigned long long)(0x20u | (vg & 0x0fu)) << 6 | ((vg_r & 7u) << 3 | (vg_b & 7u)); n = 12;               // QOI_OP_LUMA0
    } else if (q10_fits(vg_r, 8) && q10_fits(vg, 16) && q10_fits(vg_b, 8)) {
        c = (unsigned long long)(vg & 0x1fu) << 8 | (vg_r & 15u) << 4 | (vg_b & 15u); n = 14;                          // QOI_OP_LUMA
    } else if (g == r && g == b) {
        c = (unsigned long long)0xfcu << 10 | g; n = 18;                                                              // QOI_OP_GRAY
    } else if (q10_fits(vg_r, 32) && q10_fits(vg, 64) && q10_fits(vg_b, 32)) {
        c = (unsigned long long)((0x6u << 7) | (vg & 0x7fu)) << 12 | (vg_r & 63u) << 6 | (vg_b & 63u); n = 22;        // QOI_OP_LUMA2
    } else if (q10_fits(vg_r, 128) && q10_fits(vg, 256) && q10_fits(vg_b, 128)) {
        c = (unsigned long long)((0x1cu << 9) | (vg & 0x1ffu)) << 16 | (vg_r & 255u) << 8 | (vg_b & 255u); n = 30;    // QOI_OP_LUMA3
    } else {
        c = (unsigned long long)0xfdu << 30 | (unsigned long long)r << 20 | (unsigned long long)g << 10 | b; n = 38;  // QOI_OP_RGB
    }
    e.code = e.code << n | c; e.nbits += n;
    return e;
}
__device__ __forceinline__ Q10Eval q10_eval(const QeImage& im, uint32_t i, uint32_t y, uint32_t x)
{
    const unsigned long long cur = q10_load(im, y, x);
    unsigned long long prev = Q10_INITIAL;
    if (i) prev = x ? q10_load(im, y, x - 1) : q10_load(im, y - 1, im.w - 1);
    return q10_eval_px(cur, prev, y ? q10_load(im, y - 1, x) : 0ull, y > 0);
}

// ---- per tile: the last pixel that differs from its predecessor ------------------------------------------------------
__global__ void __launch_bounds__(QE_THREADS)
q10_tile_ne_kernel(const QeImage* __restrict__ imgs, QeTile* __restrict__ tiles)
{
    __shared__ int s_warp[QE_THREADS / 32];
    const QeImage& im = imgs[blockIdx.y];
    if (blockIdx.x >= im.ntiles) return;
    const uint32_t i0 = blockIdx.x * QE_TILE + threadIdx.x * QE_PER;
    int last = -1;
    if (i0 < im.np) {
        unsigned long long prev = i0 ? q10_load_i(im, i0 - 1) : Q10_INITIAL;
#pragma unroll
        for (int q = 0; q < QE_PER; ++q) {
            const uint32_t i = i0 + q;
            if (i < im.np) {
                const unsigned long long cur = q10_load_i(im, i);
                if (cur != prev) last = (int)i;
                prev = cur;
            }
        }
    }
    int tot;
    qe_cta_scan<true>(last, -1, s_warp, &tot);
    if (threadIdx.x == 0) tiles[im.tile_base + blockIdx.x].last_ne = tot;
}

// n <= 28 bits of `code`, MSB first, at bit p of s_bits (bit 0 = MSB of word 0)
__device__ __forceinline__ void q10_put(uint32_t* s_bits, uint32_t p, uint32_t code, int n)
{
    const uint32_t w = p >> 5, sh = p & 31u;
    const unsigned long long v = (unsigned long long)code << (64 - n - (int)sh);
    atomicOr(&s_bits[w], (uint32_t)(v >> 32));
    if ((uint32_t)v) atomicOr(&s_bits[w + 1], (uint32_t)v);
}

// ---- codes of a tile. EMIT = false: bits of the tile. EMIT = true: the bits, MSB first, at their place ---------------
template <bool EMIT>
__global__ void __launch_bounds__(QE_THREADS)
q10_tile_kernel(const QeImage* __restrict__ imgs, QeTile* __restrict__ tiles)
{
    __shared__ int s_warp[QE_THREADS / 32];
    __shared__ uint32_t s_bits[EMIT ? (QE_TILE * Q10_MAX_BITS / 32 + 4) : 1];
    const QeImage& im = imgs[blockIdx.y];
    if (blockIdx.x >= im.ntiles) return;
    const uint32_t tile_index = im.tile_base + blockIdx.x;
    const QeTile tile = tiles[tile_index];
    const uint32_t i0 = blockIdx.x * QE_TILE + threadIdx.x * QE_PER;
    Q10Eval ev[QE_PER + 1];
    int my_last = -1;
    {
        uint32_t y = i0 < im.np ? i0 / im.w : 0, x = i0 < im.np ? i0 - y * im.w : 0;
#pragma unroll
        for (int q = 0; q <= QE_PER; ++q) {
            const uint32_t i = i0 + q;
            ev[q].eq = false; ev[q].code = 0; ev[q].nbits = 0;
            if (i < im.np) {
                ev[q] = q10_eval(im, i, y, x);
                if (q < QE_PER && !ev[q].eq) my_last = (int)i;
                if (++x == im.w) { x = 0; ++y; }
            }
        }
    }
    int last_ne = max(tile.carry_ne, qe_cta_scan<true>(my_last, -1, s_warp, nullptr));
    unsigned long long codes[QE_PER]; int nb[QE_PER]; int mybits = 0;
#pragma unroll
    for (int q = 0; q < QE_PER; ++q) {
        const uint32_t i = i0 + q;
        codes[q] = 0; nb[q] = 0;
        if (i < im.np) {
            if (!ev[q].eq) { last_ne = (int)i; codes[q] = ev[q].code; nb[q] = ev[q].nbits; }
            else {
                // one code at the last pixel of a run of at most 256 (:238-252, :306-316)
                const uint32_t r = (i - (uint32_t)(last_ne + 1)) & 255u;             // run - 1
                if (r == 255u || i + 1 == im.np || !ev[q + 1].eq) {
                    if (r < 7u) { codes[q] = 0xf0u | r; nb[q] = 8; }
                    else { codes[q] = (0xf7u << 8) | (r - 7u); nb[q] = 16; }
                }
            }
            mybits += nb[q];
        }
    }
    int total;
    const int ex = qe_cta_scan<false>(mybits, 0, s_warp, &total);
    if (!EMIT) { if (threadIdx.x == 0) tiles[tile_index].bits = (uint32_t)total; return; }
    // as in qe_tile_kernel: the tile's bits at the bit alignment they have in memory, then big-endian words
    const uint32_t g0 = QOIX_HEADER_SIZE * 8 + tile.bit_base;
    const uint32_t mis = g0 & 31u;
    const uint32_t nwords = (mis + (uint32_t)total + 31u) >> 5;
    for (uint32_t w = threadIdx.x; w < nwords; w += QE_THREADS) s_bits[w] = 0;
    __syncthreads();
    uint32_t p = mis + (uint32_t)ex;
#pragma unroll
    for (int q = 0; q < QE_PER; ++q) {
        if (nb[q] > 26) {
            q10_put(s_bits, p, (uint32_t)(codes[q] >> 26), nb[q] - 26);
            q10_put(s_bits, p + (uint32_t)(nb[q] - 26), (uint32_t)(codes[q] & 0x3ffffffull), 26);
        } else if (nb[q]) q10_put(s_bits, p, (uint32_t)codes[q], nb[q]);
        p += (uint32_t)nb[q];
    }
    __syncthreads();
    uint32_t* const words = (uint32_t*)im.out + (g0 >> 5);
    const bool tail_shared = ((mis + (uint32_t)total) & 31u) != 0;
    for (uint32_t w = threadIdx.x; w < nwords; w += QE_THREADS) {
        const uint32_t v = __byte_perm(s_bits[w], 0, 0x0123);          // stream order = big-endian words
        if (w == 0 || (w == nwords - 1 && tail_shared)) { if (v) atomicOr(words + w, v); }
        else words[w] = v;
    }
}

// ---- host side of the image table (shared with the emulation harness) ----------------------------------------------
// qoi10b_encode's own checks (:138-146) for the channel counts routed here, plus the bound that keeps bit positions in 32 bits
inline bool q10_valid(uint32_t width, uint32_t height, int channels, int bitdepth, int compression)
{
    return (channels == 3 || channels == 4) && width && height && height < 400000000u / width && compression == 0 && bitdepth == 10 &&
           (unsigned long long)width * height * Q10_MAX_BITS + 4096 < 0xffffffffull;
}
inline bool q10_setup(QeImage& Q, const uint8_t* pixels, uint32_t width, uint32_t height, int pitch, int channels, int bitdepth,
                      int colorspace, int compression, float pixelAspectRatio, float resolutionY, uint8_t* out, uint32_t& total_tiles)
{
    if (!q10_valid(width, height, channels, bitdepth, compression) || !pixels || !out || ((uintptr_t)out & 15)) return false;
    if (((uintptr_t)pixels & 1) || (pitch & 1) || pitch < (int)(width * (uint32_t)channels * 2u)) return false;
    Q = QeImage();
    Q.pixels = pixels; Q.pitch = pitch; Q.w = width; Q.h = height; Q.np = width * height; Q.channels = channels;
    Q.marker_bits = 40u;                                                       // five 0xFF bytes (:488-491), then 1-bits to the byte boundary (:493-494)
    Q.tile_base = total_tiles; Q.ntiles = (Q.np + QE_TILE - 1) / QE_TILE; total_tiles += Q.ntiles;
    Q.out = out;
    uint8_t* h = Q.header;
    uint32_t fa, fr;
    __builtin_memcpy(&fa, &pixelAspectRatio, 4); __builtin_memcpy(&fr, &resolutionY, 4);
    const uint32_t w3[3] = {0x716F6978u, width, height}, w2[2] = {fa, fr};    // "qoix", big-endian
    for (int k = 0; k < 3; ++k) { h[4 * k] = (uint8_t)(w3[k] >> 24); h[4 * k + 1] = (uint8_t)(w3[k] >> 16); h[4 * k + 2] = (uint8_t)(w3[k] >> 8); h[4 * k + 3] = (uint8_t)w3[k]; }
    h[12] = 1;                                                                 // qoix_version (:168)
    h[13] = (uint8_t)channels; h[14] = (uint8_t)bitdepth; h[15] = (uint8_t)colorspace; h[16] = 0;
    for (int k = 0; k < 2; ++k) { h[17 + 4 * k] = (uint8_t)(w2[k] >> 24); h[18 + 4 * k] = (uint8_t)(w2[k] >> 16); h[19 + 4 * k] = (uint8_t)(w2[k] >> 8); h[20 + 4 * k] = (uint8_t)w2[k]; }
    return true;
}

}  // namespace
