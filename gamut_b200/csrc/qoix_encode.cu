// qoix_encode.cu -- QOI-Plane10 encoder on the GPU (the other side of the config-5 path).
//
// Reference: qoiplane10_encode (codecs/qoiplane10.d:99-314) as qoix_lz4_encode calls it for 10-bit 1/2-channel images
// (plugins/qoix.d:251-339). The encoder sees every pixel at once, so -- unlike the decoder -- nothing in it is serial
// except where a code lands in the bit stream:
//   * the predictor of a pixel (locoPredict of the ORIGINAL neighbours, :84-96), its residual and its alpha
//     difference depend on four input pixels only;
//   * a pixel equal to its predecessor joins a run; runs are cut every 256 pixels from the start of the maximal
//     sequence of such pixels, and a run emits ONE code at its last pixel (a run of one whose residual is tiny
//     becomes a DIFF1). The position of a pixel inside its sequence is its distance to the last pixel that differs
//     from its predecessor: a prefix maximum;
//   * the bit position of a code is the prefix sum of the code lengths before it.
// Five kernels over tiles of 1024 pixels: last differing pixel per tile, prefix maximum over the tiles of an image,
// bits per tile, prefix sum over the tiles (+ header, end marker, length), emit. Output is bit-identical to the
// reference encoder (tests/test_qoix_encode_gpu.py compares with the oracle's restatement of it and decodes the
// result with both decoders). The LZ4 stage of qoix_lz4_encode (LZ4_compress, lz4.d:329) is not built: the stream
// is returned with compression = 0, which is what the reference itself returns whenever LZ4 does not pay.
#include "../../include/gamut_b200.h"
#include "common.h"
#include <algorithm>
#include <vector>
#include <cstring>

namespace {

constexpr int QE_TILE = 1024, QE_THREADS = 256, QE_PER = QE_TILE / QE_THREADS;
constexpr int QOIX_HEADER_SIZE = 25;

struct QeImage {
    const uint8_t* pixels; int pitch;        // la16 / l16 rows
    uint32_t w, h, np; int channels;
    uint32_t tile_base, ntiles;
    uint8_t* out;                            // 25-byte header + payload
    uint32_t out_cap;
    uint8_t header[QOIX_HEADER_SIZE];
};
struct QeTile { int last_ne; int carry_ne; uint32_t bits; uint32_t bit_base; };

struct QePx { uint32_t l, a; };
__device__ __forceinline__ QePx qe_load(const QeImage& im, uint32_t y, uint32_t x)
{
    const uint16_t* p = (const uint16_t*)(im.pixels + (size_t)im.pitch * y) + (size_t)x * im.channels;
    QePx r; r.l = (uint32_t)p[0] >> 6; r.a = im.channels == 2 ? (uint32_t)p[1] >> 6 : 1023u;
    return r;
}
__device__ __forceinline__ QePx qe_load_i(const QeImage& im, uint32_t i) { const uint32_t y = i / im.w; return qe_load(im, y, i - y * im.w); }

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
__device__ __forceinline__ QeEval qe_eval(const QeImage& im, uint32_t i, uint32_t y, uint32_t x)
{
    const QePx cur = qe_load(im, y, x);
    QePx prev; prev.l = 0; prev.a = 1023;                       // initialPredictor (:59)
    if (i) prev = x ? qe_load(im, y, x - 1) : qe_load(im, y - 1, im.w - 1);
    int pred;
    if (y == 0) pred = (int)prev.l;
    else if (x == 0) pred = (int)qe_load(im, y - 1, 0).l;
    else pred = qe_med((int)prev.l, (int)qe_load(im, y - 1, x).l, (int)qe_load(im, y - 1, x - 1).l);
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
__device__ __forceinline__ void qe_code(const QeEval& e, uint32_t i, int last_ne, bool next_eq, uint32_t np, uint32_t& code, int& nbits)
{
    if (!e.eq) { code = e.code; nbits = e.nbits; return; }
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
        QePx prev; prev.l = 0; prev.a = 1023;
        if (i0) prev = qe_load_i(im, i0 - 1);
#pragma unroll
        for (int q = 0; q < QE_PER; ++q) {
            const uint32_t i = i0 + q;
            if (i < im.np) {
                const QePx cur = qe_load(im, y, x);
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
// (5 x 0xFF and 1-bits up to the byte boundary, :305-310) and the stream length; zeroes the words that two tiles share.
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
        const uint32_t pad = (8u - (E & 7u)) & 7u;
        const uint32_t total = E + 40u + pad;
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
template <bool EMIT>
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
        const bool fast = QE_PER == 4 && im.channels == 2 && i0 < im.np && y > 0 && x >= 4 && x + 8 <= im.w && (x & 3) == 0 &&
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
                    ev[q] = qe_eval(im, i, y, x);
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
            qe_code(ev[q], i, last_ne, ev[q + 1].eq, im.np, codes[q], nb[q]);
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

}  // namespace

namespace gb {

static bool qe_valid(const gb200_qoix_desc& d)
{
    // qoiplane10_encode's own checks (:101-110), plus the bound that keeps bit positions in 32 bits
    return (d.channels == 1 || d.channels == 2) && d.width && d.height && d.height < 400000000u / d.width &&
           d.compression == 0 && d.bitdepth == 10 && (unsigned long long)d.width * d.height * 28ull + 4096 < 0xffffffffull;
}

// Encodes n device-resident images into n device buffers (each at least gb200_qoix_encode_bound bytes, 16-byte
// aligned). out_len[i] = stream length, 0 for an image the encoder refuses.
bool qoiplane10_encode_device(int n, const uint8_t* const* pixels_dev, const gb200_qoix_desc* descs, uint8_t* const* out_dev,
                              int* out_len, cudaStream_t st)
{
    if (!ensure_device()) return false;
    std::vector<QeImage> imgs; std::vector<int> which;
    uint32_t total_tiles = 0;
    for (int i = 0; i < n; ++i) {
        out_len[i] = 0;
        const gb200_qoix_desc& d = descs[i];
        if (!qe_valid(d) || !pixels_dev[i] || !out_dev[i] || ((uintptr_t)out_dev[i] & 15) || ((uintptr_t)pixels_dev[i] & 1) || (d.pitchBytes & 1)) continue;
        QeImage Q; memset(&Q, 0, sizeof(Q));
        Q.pixels = pixels_dev[i]; Q.pitch = d.pitchBytes; Q.w = d.width; Q.h = d.height; Q.np = d.width * d.height; Q.channels = d.channels;
        Q.tile_base = total_tiles; Q.ntiles = (Q.np + QE_TILE - 1) / QE_TILE; total_tiles += Q.ntiles;
        Q.out = out_dev[i];
        uint8_t* h = Q.header;
        auto be32 = [&](int at, uint32_t v) { h[at] = (uint8_t)(v >> 24); h[at + 1] = (uint8_t)(v >> 16); h[at + 2] = (uint8_t)(v >> 8); h[at + 3] = (uint8_t)v; };
        be32(0, 0x716F6978u); be32(4, d.width); be32(8, d.height);
        h[12] = 2; h[13] = d.channels; h[14] = d.bitdepth; h[15] = d.colorspace; h[16] = 0;
        uint32_t f; memcpy(&f, &d.pixelAspectRatio, 4); be32(17, f); memcpy(&f, &d.resolutionY, 4); be32(21, f);
        imgs.push_back(Q); which.push_back(i);
    }
    const int m = (int)imgs.size();
    if (!m) return true;
    DevBuf d_imgs(sizeof(QeImage) * (size_t)m), d_tiles(sizeof(QeTile) * ((size_t)total_tiles + 1)), d_len(sizeof(int) * (size_t)m);
    PinnedBuf h_len(sizeof(int) * (size_t)m);
    if (!d_imgs.p || !d_tiles.p || !d_len.p || !h_len.p) return false;
    bool ok = cuda_ok(cudaMemcpyAsync(d_imgs.p, imgs.data(), sizeof(QeImage) * (size_t)m, cudaMemcpyHostToDevice, st), "qe imgs", __FILE__, __LINE__);
    if (ok) {
        uint32_t most = 0;
        for (const QeImage& Q : imgs) most = std::max(most, Q.ntiles);
        for (int k0 = 0; ok && k0 < m; k0 += 65535) {           // grid.y is limited to 65535
            const int mk = std::min(65535, m - k0);
            const dim3 grid(most, (unsigned)mk);
            const QeImage* dI = d_imgs.as<QeImage>() + k0; QeTile* dT = d_tiles.as<QeTile>(); int* dl = d_len.as<int>() + k0;
            qe_tile_ne_kernel<<<grid, QE_THREADS, 0, st>>>(dI, mk, dT);
            qe_scan_kernel<<<mk, QE_THREADS, 0, st>>>(dI, dT, 0, dl);
            qe_tile_kernel<false><<<grid, QE_THREADS, 0, st>>>(dI, mk, dT);
            qe_scan_kernel<<<mk, QE_THREADS, 0, st>>>(dI, dT, 1, dl);
            qe_tile_kernel<true><<<grid, QE_THREADS, 0, st>>>(dI, mk, dT);
            count_launch(5);
        }
        ok = ok && dev_read_back_async(h_len.p, d_len.p, sizeof(int) * (size_t)m, st);
    }
    ok = cuda_ok(cudaStreamSynchronize(st), "qe sync", __FILE__, __LINE__) && ok;
    ok = ok && cuda_ok(cudaGetLastError(), "qe kernels", __FILE__, __LINE__);
    if (ok) for (int k = 0; k < m; ++k) out_len[which[k]] = ((const int*)h_len.p)[k];
    return ok;
}

}  // namespace gb

// worst case of qoiplane10_encode's own allocation (:112-116) rounded up for the word stores of the emit kernel
GB_API size_t gb200_qoix_encode_bound(const gb200_qoix_desc* desc)
{
    if (!desc) return 0;
    const unsigned long long np = (unsigned long long)desc->width * desc->height;
    return (size_t)((np * (desc->channels == 1 ? 14 : 28) + 7) / 8 + QOIX_HEADER_SIZE + 5 + 64);
}

GB_API int gb200_qoix_encode_batch_device(int n, const uint8_t* const* pixels_dev, const gb200_qoix_desc* descs,
                                          uint8_t* const* out_dev, int* out_len, void* stream)
{
    gb::clear_error();
    if (n < 0 || !pixels_dev || !descs || !out_dev || !out_len) { gb::set_error("qoix_encode_batch_device: bad arguments"); return 0; }
    return gb::qoiplane10_encode_device(n, pixels_dev, descs, out_dev, out_len, (cudaStream_t)stream) ? 1 : 0;
}

// qoix_lz4_encode (plugins/qoix.d:251) for the images QOI-Plane10 takes (10-bit, 1 or 2 channels): host pixels in,
// malloc()'d stream out (free with gb200_free), *out_len its length. The stream is not LZ4-wrapped (compression 0).
GB_API uint8_t* gb200_qoix_encode(const uint8_t* pixels, const gb200_qoix_desc* desc, int* out_len)
{
    gb::clear_error();
    if (!gb::ensure_device()) return nullptr;
    if (!pixels || !desc || !out_len || !gb::qe_valid(*desc) || desc->pitchBytes < (int)(desc->width * desc->channels * 2)) {
        gb::set_error("qoix_encode: unsupported image (QOI-Plane10 takes 10-bit images with 1 or 2 channels)");
        return nullptr;
    }
    cudaStream_t st = gb::thread_stream();
    const size_t in_bytes = (size_t)desc->pitchBytes * desc->height, cap = gb200_qoix_encode_bound(desc);
    gb::DevBuf d_in(in_bytes), d_out(cap);
    if (!d_in.p || !d_out.p) return nullptr;
    if (!gb::cuda_ok(cudaMemcpyAsync(d_in.p, pixels, in_bytes, cudaMemcpyHostToDevice, st), "qe h2d", __FILE__, __LINE__)) { cudaStreamSynchronize(st); return nullptr; }
    const uint8_t* pin[1] = {d_in.as<uint8_t>()}; uint8_t* pout[1] = {d_out.as<uint8_t>()};
    int len = 0;
    if (!gb::qoiplane10_encode_device(1, pin, desc, pout, &len, st) || len <= 0) { cudaStreamSynchronize(st); return nullptr; }
    uint8_t* out = (uint8_t*)malloc((size_t)len);
    if (!out) return nullptr;
    const bool ok = gb::cuda_ok(cudaMemcpyAsync(out, d_out.p, (size_t)len, cudaMemcpyDeviceToHost, st), "qe d2h", __FILE__, __LINE__) &&
                    gb::cuda_ok(cudaStreamSynchronize(st), "qe sync", __FILE__, __LINE__);
    if (!ok) { cudaStreamSynchronize(st); free(out); return nullptr; }
    *out_len = len;
    return out;
}
