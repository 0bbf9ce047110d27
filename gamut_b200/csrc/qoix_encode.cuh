// qoix_encode.cuh -- the kernels of qoix_encode.cu (see the comment at the top of that file for the formulation).
// Kept apart from the host code so that tests/test_qoix_encode_emulated.py can compile exactly this text for the host
// under a thread-per-CUDA-thread emulation (tests/cuda_emu.h) and compare it with the oracle without a GPU.
#pragma once
#include <stdint.h>
#include <stddef.h>

namespace {

constexpr int QE_TILE = 1024, QE_THREADS = 256, QE_PER = QE_TILE / QE_THREADS;
constexpr int QOIX_HEADER_SIZE = 25;

struct QeImage {
    const uint8_t* pixels; int pitch;        // la16 / l16 rows (QOI-Plane10) or la8 / l8 rows (QOI-Plane)
    uint32_t w, h, np; int channels;
    uint32_t marker_bits;                    // 1-bits after the last code, before the 1-bits that fill the last byte
    uint32_t tile_base, ntiles;
    uint8_t* out;                            // 25-byte header + payload
    uint32_t out_cap;
    uint8_t header[QOIX_HEADER_SIZE];
};
struct QeTile { int last_ne; int carry_ne; uint32_t bits; uint32_t bit_base; };

// P8 = false: QOI-Plane10 (qoiplane10.d), 10-bit samples in the top bits of 16-bit words. P8 = true: QOI-Plane
// (qoiplane.d:109-375), 8-bit samples. The two encoders share everything but the codes themselves.
struct QePx { uint32_t l, a; };
template <bool P8>
__device__ __forceinline__ QePx qe_load(const QeImage& im, uint32_t y, uint32_t x)
{
    QePx r;
    if (P8) {
        const uint8_t* p = im.pixels + (size_t)im.pitch * y + (size_t)x * im.channels;
        r.l = p[0]; r.a = im.channels == 2 ? (uint32_t)p[1] : 255u;
    } else {
        const uint16_t* p = (const uint16_t*)(im.pixels + (size_t)im.pitch * y) + (size_t)x * im.channels;
        r.l = (uint32_t)p[0] >> 6; r.a = im.channels == 2 ? (uint32_t)p[1] >> 6 : 1023u;
    }
    return r;
}
template <bool P8>
__device__ __forceinline__ QePx qe_load_i(const QeImage& im, uint32_t i) { const uint32_t y = i / im.w; return qe_load<P8>(im, y, i - y * im.w); }

__device__ __forceinline__ int qe_med(int left, int top, int topleft)       // locoPredict, qoiplane10.d:84-96
{
    const int mx = max(left, top), mn = min(left, top);
    if (topleft >= mx) return mn;
    if (topleft <= mn) return mx;
    return min(max(left + top - topleft, 0), 1023);
}

// Everything about pixel i that does not depend on other tiles: the pixel, whether it equals its predecessor, and the
// code it would emit as a pixel of its own (the DIFF / ADIFF / LA part of the encoder's loop body, :230-262).
struct QeEval { bool eq; uint32_t code; int nbits; uint32_t diff1; bool diff1_ok; };
__device__ __forceinline__ QeEval qe_eval_px(QePx cur, QePx prev, int pred)
{
    QeEval e;
    e.eq = cur.l == prev.l && cur.a == prev.a;
    const uint32_t vg = (cur.l - (uint32_t)pred) & 1023u;
    e.diff1 = vg & 7u; e.diff1_ok = vg < 4 || vg >= 1024 - 4;
    e.code = 0; e.nbits = 0;
    if (!e.eq) {
        const uint32_t va = (cur.a - prev.a) & 1023u;
        if (va) {
            if (va < 32 || va >= 1024 - 32) { e.code = (0x3eu << 6) | (va & 0x3fu); e.nbits = 12; }
            else { e.code = (0xfeu << 20) | (cur.l << 10) | cur.a; e.nbits = 28; return e; }
        }
        if (e.diff1_ok) { e.code = (e.code << 4) | e.diff1; e.nbits += 4; }
        else if (vg < 32 || vg >= 1024 - 32) { e.code = (e.code << 8) | 0x80u | (vg & 0x3fu); e.nbits += 8; }
        else if (vg < 64 || vg >= 1024 - 64) { e.code = (e.code << 12) | (0x1eu << 7) | (vg & 0x7fu); e.nbits += 12; }
        else { e.code = (e.code << 14) | (0xeu << 10) | vg; e.nbits += 14; }
    }
    return e;
}
// QOI-Plane (qoiplane.d:250-311): nibble-aligned codes; `top` = the pixel above, or the previous pixel in the first row
__device__ __forceinline__ QeEval qe_eval_px8(QePx cur, QePx prev, uint32_t top)
{
    QeEval e;
    e.eq = cur.l == prev.l && cur.a == prev.a;
    e.diff1 = 0; e.diff1_ok = false; e.code = 0; e.nbits = 0;
    if (!e.eq) {
        const int va = (int)(signed char)(cur.a - prev.a);
        if (va) {
            if (va >= -7 && va <= 7) { e.code = 0xb0u | (uint32_t)(va + 8); e.nbits = 8; }                 // QOIPLANE_ADIFF
            else { e.code = (0xb0u << 16) | (cur.l << 8) | cur.a; e.nbits = 24; return e; }                // QOIPLANE_LA
        }
        const uint32_t avg = (top + prev.l + 1u) >> 1;
        const int d = (int)(signed char)(cur.l - avg);
        if (d >= -4 && d <= 3) { e.code = (e.code << 4) | (uint32_t)(d + 4); e.nbits += 4; }               // QOIPLANE_DIFF1
        else if (d >= -16 && d <= 15) { e.code = (e.code << 8) | 0x80u | (uint32_t)(d + 16); e.nbits += 8; } // QOIPLANE_DIFF2
        else { e.code = (e.code << 12) | 0xa00u | cur.l; e.nbits += 12; }                                  // QOIPLANE_DIRECT
    }
    return e;
}
template <bool P8>
__device__ __forceinline__ QeEval qe_eval(const QeImage& im, uint32_t i, uint32_t y, uint32_t x)
{
    const QePx cur = qe_load<P8>(im, y, x);
    QePx prev; prev.l = 0; prev.a = P8 ? 255 : 1023;            // initialPredictor (qoiplane10.d:59, qoiplane.d:95)
    if (i) prev = x ? qe_load<P8>(im, y, x - 1) : qe_load<P8>(im, y - 1, im.w - 1);
    if (P8) return qe_eval_px8(cur, prev, y ? qe_load<P8>(im, y - 1, x).l : prev.l);
    int pred;
    if (y == 0) pred = (int)prev.l;
    else if (x == 0) pred = (int)qe_load<P8>(im, y - 1, 0).l;
    else pred = qe_med((int)prev.l, (int)qe_load<P8>(im, y - 1, x).l, (int)qe_load<P8>(im, y - 1, x - 1).l);
    return qe_eval_px(cur, prev, pred);
}
// la16 pixels x0-1 .. x0+4 of row y (x0 a multiple of 4 inside the row, the row 16-byte aligned): one vector + two pixels
__device__ __forceinline__ void qe_load6_la(const QeImage& im, uint32_t y, uint32_t x0, QePx (&p)[QE_PER + 2])
{
    const uint32_t* row = (const uint32_t*)(im.pixels + (size_t)im.pitch * y);
    const uint4 v = __ldg((const uint4*)(row + x0));
    const uint32_t w[6] = {__ldg(row + x0 - 1), v.x, v.y, v.z, v.w, __ldg(row + x0 + 4)};
#pragma unroll
    for (int k = 0; k < 6; ++k) { p[k].l = (w[k] & 0xffffu) >> 6; p[k].a = w[k] >> 22; }
}

// the code a pixel emits given its place in its run: nothing inside a run, the run's code at its last pixel
template <bool P8>
__device__ __forceinline__ void qe_code(const QeEval& e, uint32_t i, int last_ne, bool next_eq, uint32_t np, uint32_t& code, int& nbits)
{
    if (!e.eq) { code = e.code; nbits = e.nbits; return; }
    if (P8) {
        const uint32_t r8 = (i - (uint32_t)(last_ne + 1)) % 258u;  // index inside the run of at most 258 (qoiplane.d:236-241)
        code = 0; nbits = 0;
        if (!(r8 == 257u || i + 1 == np || !next_eq)) return;
        if (r8 < 3) { code = 0xcu | r8; nbits = 4; }               // QOIPLANE_REPEAT1: run - 1 = r8 (:181-185)
        else { code = 0xf00u | (r8 - 3u); nbits = 12; }            // QOIPLANE_REPEAT2: run - 4 (:194-197)
        return;
    }
    const uint32_t r = (i - (uint32_t)(last_ne + 1)) & 255u;   // index inside the run of at most 256 (:224-228)
    const bool end = r == 255u || i + 1 == np || !next_eq;
    code = 0; nbits = 0;
    if (!end) return;
    if (r == 0 && e.diff1_ok) { code = e.diff1; nbits = 4; return; }     // FLUSH_RUN with run == 1
    if (r < 7) { code = 0x30u | r; nbits = 6; }                            // ENCODE_RUN: run - 1 = r
    else { code = (0x37u << 8) | (r - 7u); nbits = 14; }
}

// inclusive prefix over the CTA (max or sum) of one value per thread; returns the exclusive value, *total = all
template <bool MAX>
__device__ __forceinline__ int qe_cta_scan(int v, int identity, int* s_warp, int* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int n = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc = MAX ? max(inc, n) : inc + n; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int off = identity, tot = identity;
#pragma unroll
    for (int w = 0; w < QE_THREADS / 32; ++w) { const int c = s_warp[w]; if (w < warp) off = MAX ? max(off, c) : off + c; tot = MAX ? max(tot, c) : tot + c; }
    int ex = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) ex = identity;
    __syncthreads();
    if (total) *total = tot;
    return MAX ? max(off, ex) : off + ex;
}

// ---- E1: index of the last pixel of the tile that differs from its predecessor (-1: none) ------------------------
template <bool P8>
__global__ void __launch_bounds__(QE_THREADS)
qe_tile_ne_kernel(const QeImage* __restrict__ imgs, int nimgs, QeTile* __restrict__ tiles)
{
    __shared__ int s_warp[QE_THREADS / 32];
    // grid = (most tiles of an image, images): no search for the image at the start of every CTA
    const QeImage& im = imgs[blockIdx.y];
    if (blockIdx.x >= im.ntiles) return;
    const uint32_t tile_index = im.tile_base + blockIdx.x;
    const uint32_t i0 = blockIdx.x * QE_TILE + threadIdx.x * QE_PER;
    int last = -1;
    if (i0 < im.np) {
        uint32_t y = i0 / im.w, x = i0 - y * im.w;
        QePx prev; prev.l = 0; prev.a = P8 ? 255 : 1023;
        if (i0) prev = qe_load_i<P8>(im, i0 - 1);
#pragma unroll
        for (int q = 0; q < QE_PER; ++q) {
            const uint32_t i = i0 + q;
            if (i < im.np) {
                const QePx cur = qe_load<P8>(im, y, x);
                if (cur.l != prev.l || cur.a != prev.a) last = (int)i;
                prev = cur;
                if (++x == im.w) { x = 0; ++y; }
            }
        }
    }
    int tot;
    qe_cta_scan<true>(last, -1, s_warp, &tot);
    if (threadIdx.x == 0) tiles[tile_index].last_ne = tot;
}

// ---- E2 / E4: per image, exclusive prefix over its tiles (one CTA per image) ---------------------------------------
// phase 0: prefix maximum of last_ne -> carry_ne. phase 1: prefix sum of bits -> bit_base, then header, end marker
// (marker_bits 1-bits, then 1-bits up to the byte boundary: 5 x 0xFF for QOI-Plane10, qoiplane10.d:305-310; nine 0xF
// nibbles for QOI-Plane, qoiplane.d:318-322) and the stream length; zeroes the words that two tiles share.
__global__ void __launch_bounds__(QE_THREADS)
qe_scan_kernel(const QeImage* __restrict__ imgs, QeTile* __restrict__ tiles, int phase, int* __restrict__ out_len)
{
    __shared__ int s_warp[QE_THREADS / 32];
    const QeImage& im = imgs[blockIdx.x];
    QeTile* T = tiles + im.tile_base;
    if (phase == 0) {
        int carry = -1;
        for (uint32_t t0 = 0; t0 < im.ntiles; t0 += QE_THREADS) {
            const uint32_t t = t0 + threadIdx.x;
            const int v = t < im.ntiles ? T[t].last_ne : -1;
            int tot;
            const int ex = qe_cta_scan<true>(v, -1, s_warp, &tot);
            if (t < im.ntiles) T[t].carry_ne = max(carry, ex);
            carry = max(carry, tot);
        }
        return;
    }
    // bit positions are relative to the payload (byte 25 of the stream); a stream holds fewer than 2^32 bits only
    // for np < 153e6 pixels of 28 bits: the host refuses larger images
    uint32_t carry = 0;
    uint32_t* const words = (uint32_t*)im.out;                   // out is 16-byte aligned
    for (uint32_t t0 = 0; t0 < im.ntiles; t0 += QE_THREADS) {
        const uint32_t t = t0 + threadIdx.x;
        const int v = t < im.ntiles ? (int)T[t].bits : 0;
        int tot;
        const int ex = qe_cta_scan<false>(v, 0, s_warp, &tot);
        if (t < im.ntiles) {
            const uint32_t bb = carry + (uint32_t)ex;
            T[t].bit_base = bb;
            words[(QOIX_HEADER_SIZE * 8 + bb) >> 5] = 0;         // the word a tile starts in may be shared with the tile before it
        }
        carry += (uint32_t)tot;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t E = carry;                                // end of the pixel codes
        const uint32_t pad = (8u - ((E + im.marker_bits) & 7u)) & 7u;
        const uint32_t total = E + im.marker_bits + pad;
        uint8_t* o = im.out;
        // the three words the end marker touches (after the zeroing above, before any tile ORs its bits in)
        const uint32_t wfirst = (QOIX_HEADER_SIZE * 8 + E) >> 5, wlast = (QOIX_HEADER_SIZE * 8 + total - 1) >> 5;
        for (uint32_t w = wfirst; w <= wlast; ++w) words[w] = 0;
        for (uint32_t b = E; b < total; ++b) { const uint32_t p = QOIX_HEADER_SIZE * 8 + b; o[p >> 3] |= (uint8_t)(0x80u >> (p & 7u)); }
        for (int k = 0; k < QOIX_HEADER_SIZE; ++k) o[k] = im.header[k];
        out_len[blockIdx.x] = QOIX_HEADER_SIZE + (int)(total >> 3);
    }
}

// ---- E3 / E5: codes of a tile. EMIT = false: bits of the tile. EMIT = true: the bits, MSB first, at their place ---
template <bool EMIT, bool P8>
__global__ void __launch_bounds__(QE_THREADS)
qe_tile_kernel(const QeImage* __restrict__ imgs, int nimgs, QeTile* __restrict__ tiles)
{
    __shared__ int s_warp[QE_THREADS / 32];
    __shared__ uint32_t s_bits[EMIT ? (QE_TILE * 28 / 32 + 4) : 1];
    const QeImage& im = imgs[blockIdx.y];
    if (blockIdx.x >= im.ntiles) return;
    const uint32_t tile_index = im.tile_base + blockIdx.x;
    const QeTile tile = tiles[tile_index];
    const uint32_t i0 = blockIdx.x * QE_TILE + threadIdx.x * QE_PER;
    // evaluate my pixels and the one after them (whose eq decides whether my last pixel ends a run)
    QeEval ev[QE_PER + 1];
    int my_last = -1;
    {
        uint32_t y = i0 < im.np ? i0 / im.w : 0, x = i0 < im.np ? i0 - y * im.w : 0;
        // interior of a row of an aligned la16 image: the six pixels of this row and of the row above as vectors
        const bool fast = !P8 && QE_PER == 4 && im.channels == 2 && i0 < im.np && y > 0 && x >= 4 && x + 8 <= im.w && (x & 3) == 0 &&
                          (im.pitch & 15) == 0 && ((uintptr_t)im.pixels & 15) == 0;
        if (fast) {
            QePx c[QE_PER + 2], u[QE_PER + 2];
            qe_load6_la(im, y, x, c); qe_load6_la(im, y - 1, x, u);
#pragma unroll
            for (int q = 0; q <= QE_PER; ++q) {
                ev[q] = qe_eval_px(c[q + 1], c[q], qe_med((int)c[q].l, (int)u[q + 1].l, (int)u[q].l));
                if (q < QE_PER && !ev[q].eq) my_last = (int)(i0 + q);
            }
        } else {
#pragma unroll
            for (int q = 0; q <= QE_PER; ++q) {
                const uint32_t i = i0 + q;
                ev[q].eq = false; ev[q].code = 0; ev[q].nbits = 0; ev[q].diff1 = 0; ev[q].diff1_ok = false;
                if (i < im.np) {
                    ev[q] = qe_eval<P8>(im, i, y, x);
                    if (q < QE_PER && !ev[q].eq) my_last = (int)i;
                    if (++x == im.w) { x = 0; ++y; }
                }
            }
        }
    }
    int last_ne = max(tile.carry_ne, qe_cta_scan<true>(my_last, -1, s_warp, nullptr));
    uint32_t codes[QE_PER]; int nb[QE_PER]; int mybits = 0;
#pragma unroll
    for (int q = 0; q < QE_PER; ++q) {
        const uint32_t i = i0 + q;
        codes[q] = 0; nb[q] = 0;
        if (i < im.np) {
            if (!ev[q].eq) last_ne = (int)i;
            qe_code<P8>(ev[q], i, last_ne, ev[q + 1].eq, im.np, codes[q], nb[q]);
            mybits += nb[q];
        }
    }
    int total;
    const int ex = qe_cta_scan<false>(mybits, 0, s_warp, &total);
    if (!EMIT) { if (threadIdx.x == 0) tiles[tile_index].bits = (uint32_t)total; return; }
    // the tile's bits are put together in shared memory at the bit alignment they have in memory (bit 0 of s_bits =
    // the first bit of the aligned 32-bit word the tile starts in), MSB first
    const uint32_t g0 = QOIX_HEADER_SIZE * 8 + tile.bit_base;          // stream bit of the tile's first bit
    const uint32_t mis = g0 & 31u;
    const uint32_t nwords = (mis + (uint32_t)total + 31u) >> 5;
    for (uint32_t w = threadIdx.x; w < nwords; w += QE_THREADS) s_bits[w] = 0;
    __syncthreads();
    uint32_t p = mis + (uint32_t)ex;
#pragma unroll
    for (int q = 0; q < QE_PER; ++q) {
        if (nb[q]) {
            const uint32_t w = p >> 5, sh = p & 31u;
            const unsigned long long v = (unsigned long long)codes[q] << (64 - nb[q] - (int)sh);     // nbits <= 28, sh <= 31
            atomicOr(&s_bits[w], (uint32_t)(v >> 32));
            if ((uint32_t)v) atomicOr(&s_bits[w + 1], (uint32_t)v);
            p += (uint32_t)nb[q];
        }
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
// qoiplane10_encode's / qoiplane_encode's own checks (qoiplane10.d:101-110, qoiplane.d:111-120), plus the bound that
// keeps bit positions in 32 bits
inline bool qe_valid(uint32_t width, uint32_t height, int channels, int bitdepth, int compression)
{
    return (channels == 1 || channels == 2) && width && height && height < 400000000u / width && compression == 0 &&
           (bitdepth == 10 || bitdepth == 8) && (unsigned long long)width * height * 28ull + 4096 < 0xffffffffull;
}
// Fills the table entry of one image; false = the encoder refuses it. total_tiles is advanced by the image's tiles.
inline bool qe_setup(QeImage& Q, const uint8_t* pixels, uint32_t width, uint32_t height, int pitch, int channels, int bitdepth,
                     int colorspace, int compression, float pixelAspectRatio, float resolutionY, uint8_t* out, uint32_t& total_tiles)
{
    if (!qe_valid(width, height, channels, bitdepth, compression) || !pixels || !out || ((uintptr_t)out & 15)) return false;
    if (bitdepth == 10 && (((uintptr_t)pixels & 1) || (pitch & 1))) return false;
    if (pitch < (int)(width * (uint32_t)channels * (bitdepth == 10 ? 2u : 1u))) return false;
    Q = QeImage();
    Q.pixels = pixels; Q.pitch = pitch; Q.w = width; Q.h = height; Q.np = width * height; Q.channels = channels;
    Q.marker_bits = bitdepth == 10 ? 40u : 36u;
    Q.tile_base = total_tiles; Q.ntiles = (Q.np + QE_TILE - 1) / QE_TILE; total_tiles += Q.ntiles;
    Q.out = out;
    uint8_t* h = Q.header;
    uint32_t fa, fr;
    __builtin_memcpy(&fa, &pixelAspectRatio, 4); __builtin_memcpy(&fr, &resolutionY, 4);
    const uint32_t w3[3] = {0x716F6978u, width, height}, w2[2] = {fa, fr};    // "qoix", big-endian
    for (int k = 0; k < 3; ++k) { h[4 * k] = (uint8_t)(w3[k] >> 24); h[4 * k + 1] = (uint8_t)(w3[k] >> 16); h[4 * k + 2] = (uint8_t)(w3[k] >> 8); h[4 * k + 3] = (uint8_t)w3[k]; }
    h[12] = bitdepth == 10 ? 2 : 1;                                           // version_: qoiplane10.d:133, qoiplane.d:153
    h[13] = (uint8_t)channels; h[14] = (uint8_t)bitdepth; h[15] = (uint8_t)colorspace; h[16] = 0;
    for (int k = 0; k < 2; ++k) { h[17 + 4 * k] = (uint8_t)(w2[k] >> 24); h[18 + 4 * k] = (uint8_t)(w2[k] >> 16); h[19 + 4 * k] = (uint8_t)(w2[k] >> 8); h[20 + 4 * k] = (uint8_t)w2[k]; }
    return true;
}

}  // namespace
