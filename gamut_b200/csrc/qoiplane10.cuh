// qoiplane10.cuh -- parallel QOI-Plane10 decode (included by qoix.cu).
//
// qoiplane10_decode (codecs/qoiplane10.d:317-515) is two serial dependencies wound together: the opcode
// stream (2-bit aligned variable-length codes, every position depends on all codes before it) and the
// pixel recurrence (l = MED(left, up, upleft) + residual). They are taken apart:
//
//  A. parse, chunk-parallel.  Opcode lengths depend on the bits only, never on pixel values, and a parser
//     started at a wrong position falls onto true code boundaries after a few codes. The stream is cut into
//     128-byte chunks, one thread each:
//       p10_sync_kernel   pass 0 parses every chunk from its first bit; passes 1..n re-parse the chunks whose
//                         predecessor's exit position changed, until nothing changes. Chunk 0 starts at the
//                         true position, so the fixed point is the serial parse (induction over chunks).
//                         Every parse also records the chunk's pixel count, its effect on the running alpha
//                         (set v / add d) and whether it met END.
//       p10_scan_kernel   one CTA per image: exclusive scan of (pixels, alpha transform, ended) over chunks.
//       p10_write_kernel  parses once more from the true entry state and writes one 32-bit record per pixel:
//                         kind (DIFF residual / COPY of the previous pixel / literal) + residual + resolved alpha.
//  B. p10_recon_kernel, one warp per image, 32 rows per band as a skewed wavefront: lane r runs one 8-pixel
//     block behind lane r-1, so the row above arrives by warp shuffle off the critical path and the chain
//     per pixel is just the MED predictor. A run that crosses a row boundary makes the first pixel of a row
//     depend on the end of the row above; such rows start later (exactly as late as the dependency demands)
//     and read the row above from memory.
#pragma once

constexpr int P10_CHUNK_BYTES = 128;
constexpr int P10_CHUNK_BITS = P10_CHUNK_BYTES * 8;
constexpr uint32_t P10_KIND_DIFF = 0u, P10_KIND_COPY = 1u << 20, P10_KIND_LIT = 2u << 20;

struct P10Image {
    const uint8_t* stream;      // payload with its 25-byte header
    uint32_t size;              // bytes in `stream`
    uint32_t w, h, wp;          // wp = record pitch (w rounded up to 8)
    int channels;
    int image;                  // index into status[]
    uint32_t chunk_base, nchunks;
    uint32_t* recs;             // [h][wp]
    uint32_t* rowinfo;          // [h]: 0 = row starts normally, x+2 = first pixel copies column x of the row above (x = -1: any)
    uint8_t* out;
};

struct P10Chunk {               // per-chunk parse summary
    uint32_t exit_bit;          // position of the first opcode group that starts at or after the chunk's end
    uint32_t npix;              // pixels produced by the groups that start in this chunk
    uint32_t alpha;             // bit 31: ended, bit 30: set, bits 0-9: value (set) or delta (add)
};
struct P10Entry { uint32_t pix; uint32_t alpha; };     // alpha bit 31: the stream ended before this chunk

__device__ __forceinline__ uint32_t p10_find_image(const P10Image* __restrict__ imgs, int n, uint32_t c)
{
    int lo = 0, hi = n - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (imgs[mid].chunk_base <= c) lo = mid; else hi = mid - 1; }
    return (uint32_t)lo;
}

// 32 bits of the opcode stream starting at bit `bp` (MSB first). Bytes past the end read as 0xFF (== END).
struct P10Bits {
    const uint32_t* words; uint32_t adj, total_bits;
    __device__ __forceinline__ void init(const uint8_t* stream, uint32_t size)
    {
        const uintptr_t a = (uintptr_t)(stream + 25);
        words = (const uint32_t*)(a & ~(uintptr_t)3); adj = (uint32_t)(a & 3) * 8;
        total_bits = (size - 25) * 8;
    }
    __device__ __forceinline__ uint32_t peek32(uint32_t bp) const
    {
        if (bp + 64 <= total_bits) {
            const uint32_t b = bp + adj, idx = b >> 5, sh = b & 31;
            const uint32_t w0 = __byte_perm(words[idx], 0, 0x0123), w1 = __byte_perm(words[idx + 1], 0, 0x0123);
            return __funnelshift_l(w1, w0, sh);
        }
        const uint8_t* p = (const uint8_t*)words + (adj >> 3);
        uint64_t v = 0;
        const uint32_t byte0 = bp >> 3;
#pragma unroll
        for (int i = 0; i < 5; ++i) { const uint32_t bi = byte0 + i; v = (v << 8) | (bi * 8 < total_bits ? p[bi] : 0xFFu); }
        return (uint32_t)(v >> (8 - (bp & 7)));
    }
};

__device__ __forceinline__ int p10_sext(uint32_t v, int bits) { return (int)(v << (32 - bits)) >> (32 - bits); }

// Parses the opcode groups that start in [bitpos, limit). EMIT receives every group:
//   emit(kind_and_value, alpha_changed, alpha_is_literal, alpha_value_or_delta, npixels)
template <class Emit>
__device__ __forceinline__ bool p10_parse(const P10Bits& B, uint32_t& bitpos, uint32_t limit, Emit&& emit)
{
    uint32_t bp = bitpos;
    bool ended = false;
    while (bp < limit) {
        bool achg = false; int adelta = 0;
        for (;;) {
            const uint32_t v = B.peek32(bp);
            const uint32_t op = v >> 24;
            if (op < 0x80) { emit(P10_KIND_DIFF | ((uint32_t)p10_sext((op >> 4) & 7, 3) & 1023u), achg, false, adelta, 1u); bp += 4; break; }
            if (op < 0xc0) { emit(P10_KIND_DIFF | ((uint32_t)p10_sext(op & 0x3f, 6) & 1023u), achg, false, adelta, 1u); bp += 8; break; }
            if (op < 0xe0) {
                uint32_t run = (op >> 2) & 7, len = 6;
                if (run == 7) { run = ((v >> 18) & 0xff) + 7; len = 14; }
                emit(P10_KIND_COPY, achg, false, adelta, run + 1); bp += len; break;
            }
            if (op < 0xf0) { emit(P10_KIND_DIFF | ((v >> 18) & 1023u), achg, false, adelta, 1u); bp += 14; break; }
            if (op < 0xf8) { emit(P10_KIND_DIFF | ((uint32_t)p10_sext((v >> 20) & 0x7f, 7) & 1023u), achg, false, adelta, 1u); bp += 12; break; }
            if (op < 0xfc) { achg = true; adelta = p10_sext((v >> 20) & 0x3f, 6); bp += 12; continue; }
            if (op == 0xfe) { emit(P10_KIND_LIT | ((v >> 14) & 1023u), true, true, (int)((v >> 4) & 1023u), 1u); bp += 28; break; }
            ended = true; break;     // END (0xff) or the reserved 0xfc/0xfd
        }
        if (ended) break;
    }
    bitpos = bp;
    return ended;
}

// ---- A1. speculative parse + relaxation ---------------------------------------------------------------
__global__ void __launch_bounds__(128)
p10_sync_kernel(const P10Image* __restrict__ imgs, int nimgs, uint32_t total_chunks, P10Chunk* chunks,
                const uint8_t* __restrict__ dirty_in, uint8_t* __restrict__ dirty_out, int pass, uint32_t* changed)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= total_chunks) return;
    const uint32_t ii = p10_find_image(imgs, nimgs, c);
    const P10Image& im = imgs[ii];
    const uint32_t lc = c - im.chunk_base;
    if (pass > 0) {
        dirty_out[c] = 0;
        if (lc == 0 || !dirty_in[c - 1]) return;
    }
    P10Bits B; B.init(im.stream, im.size);
    const uint32_t limit = min((lc + 1) * (uint32_t)P10_CHUNK_BITS, B.total_bits);
    uint32_t bp = lc * (uint32_t)P10_CHUNK_BITS;
    if (pass > 0) bp = __ldcg(&chunks[c - 1].exit_bit);
    uint32_t npix = 0; bool aset = false; int aval = 0;
    const bool ended = p10_parse(B, bp, limit, [&](uint32_t, bool achg, bool alit, int av, uint32_t n) {
        npix += n;
        if (alit) { aset = true; aval = av; } else if (achg) aval += av;
    }) || bp >= B.total_bits;
    // A parse that stopped at END leaves a neutral exit position (the next chunk's own first bit): after a true END
    // the later chunks do not matter, and after a false one (a speculative parse on garbage) the successor keeps
    // its own guess instead of inheriting a stuck position.
    if (ended) bp = (lc + 1) * (uint32_t)P10_CHUNK_BITS;
    const uint32_t alpha = (ended ? 0x80000000u : 0u) | (aset ? 0x40000000u : 0u) | ((uint32_t)aval & 1023u);
    const uint32_t old = __ldcg(&chunks[c].exit_bit);
    __stcg(&chunks[c].exit_bit, bp); chunks[c].npix = npix; chunks[c].alpha = alpha;
    if (pass == 0) dirty_out[c] = 1;
    else if (old != bp) { dirty_out[c] = 1; atomicAdd(changed, 1u); }
}

// ---- A2. scan over the chunks of an image ----------------------------------------------------------------
__device__ __forceinline__ uint32_t p10_compose(uint32_t first, uint32_t second)   // alpha transforms, ended is sticky
{
    if (first & 0x80000000u) return first;                                         // nothing after END counts
    uint32_t r;
    if (second & 0x40000000u) r = 0x40000000u | (second & 1023u);
    else r = (first & 0x40000000u) | ((first + second) & 1023u);
    return r | (second & 0x80000000u);
}

__global__ void __launch_bounds__(256)
p10_scan_kernel(const P10Image* __restrict__ imgs, const P10Chunk* __restrict__ chunks, P10Entry* __restrict__ entries,
                uint32_t* __restrict__ ndecoded)
{
    __shared__ uint32_t s_pix[256], s_alpha[256];
    const P10Image& im = imgs[blockIdx.x];
    const int tid = threadIdx.x;
    const uint32_t per = (im.nchunks + 255) / 256;
    const uint32_t c0 = min(tid * per, im.nchunks), c1 = min(c0 + per, im.nchunks);
    uint32_t pix = 0, alpha = 0;       // identity: add 0
    for (uint32_t c = c0; c < c1; ++c) {
        const P10Chunk k = chunks[im.chunk_base + c];
        if (!(alpha & 0x80000000u)) pix += k.npix;
        alpha = p10_compose(alpha, k.alpha);
    }
    s_pix[tid] = pix; s_alpha[tid] = alpha;
    __syncthreads();
    if (tid == 0) {
        uint32_t p = 0, a = 0x40000000u | 1023u;          // initial predictor a = 1023 (qoiplane10.d:59)
        for (int i = 0; i < 256; ++i) {
            const uint32_t tp = s_pix[i], ta = s_alpha[i];
            s_pix[i] = p; s_alpha[i] = a;
            if (!(a & 0x80000000u)) p += tp;
            a = p10_compose(a, ta);
        }
        const unsigned long long np = (unsigned long long)im.w * im.h;
        ndecoded[blockIdx.x] = (uint32_t)min((unsigned long long)p, np);
    }
    __syncthreads();
    pix = s_pix[tid]; alpha = s_alpha[tid];
    for (uint32_t c = c0; c < c1; ++c) {
        const P10Chunk k = chunks[im.chunk_base + c];
        entries[im.chunk_base + c] = P10Entry{pix, alpha};
        if (!(alpha & 0x80000000u)) pix += k.npix;
        alpha = p10_compose(alpha, k.alpha);
    }
}

// ---- A3. per-pixel records -------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
p10_write_kernel(const P10Image* __restrict__ imgs, int nimgs, uint32_t total_chunks, const P10Chunk* __restrict__ chunks,
                 const P10Entry* __restrict__ entries)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= total_chunks) return;
    const uint32_t ii = p10_find_image(imgs, nimgs, c);
    const P10Image& im = imgs[ii];
    const uint32_t lc = c - im.chunk_base;
    const P10Entry e = entries[c];
    if (e.alpha & 0x80000000u) return;
    P10Bits B; B.init(im.stream, im.size);
    const uint32_t limit = min((lc + 1) * (uint32_t)P10_CHUNK_BITS, B.total_bits);
    uint32_t bp = lc ? chunks[c - 1].exit_bit : 0u;
    const unsigned long long np = (unsigned long long)im.w * im.h;
    unsigned long long i = e.pix;
    if (i >= np) return;
    uint32_t y = (uint32_t)(i / im.w), x = (uint32_t)(i - (unsigned long long)y * im.w);
    uint32_t a = e.alpha & 1023u;
    uint32_t* __restrict__ rec = im.recs + (size_t)y * im.wp + x;
    const uint32_t W = im.w, skip = im.wp - im.w;
    p10_parse(B, bp, limit, [&](uint32_t kv, bool achg, bool alit, int av, uint32_t n) {
        if (alit) a = (uint32_t)av; else if (achg) a = (a + (uint32_t)av) & 1023u;
        const uint32_t r = kv | (a << 10);
        const unsigned long long p0 = i;
        for (uint32_t q = 0; q < n && i < np; ++q, ++i) {
            if (x == 0) {
                uint32_t info = 0;
                if (y > 0 && (kv & (3u << 20)) == P10_KIND_COPY) {
                    // the pixel copies pixel p0-1 (and everything between is a copy of it too)
                    const unsigned long long dist = i - (p0 - 1);
                    info = dist <= W ? (W - (uint32_t)dist) + 2u : 1u;
                }
                im.rowinfo[y] = info;
            }
            *rec++ = r;
            if (++x == W) { x = 0; ++y; rec += skip; }
        }
    });
}

// ---- B. reconstruction -------------------------------------------------------------------------------------
__device__ __forceinline__ int p10_med(int left, int top, int topleft)       // locoPredict, qoiplane10.d:84-96
{
    const int mx = max(left, top), mn = min(left, top);
    if (topleft >= mx) return mn;
    if (topleft <= mn) return mx;
    return min(max(left + top - topleft, 0), 1023);
}

template <int CH>
__device__ __forceinline__ int p10_out_l(const uint8_t* out, uint32_t W, uint32_t y, uint32_t x)
{
    return (int)(__ldcg((const uint16_t*)out + ((size_t)y * W + x) * CH) >> 6);
}

constexpr int P10_RECON_WARPS = 8;  // warps per image: warp w owns bands w, w+8, ...; a band follows the band above
                                       // as soon as that band's last row is two blocks ahead (progress counters)
template <int CH>
__global__ void __launch_bounds__(32 * P10_RECON_WARPS)
p10_recon_kernel(const P10Image* __restrict__ imgs, const uint32_t* __restrict__ ndecoded, const int* __restrict__ status)
{
    __shared__ volatile uint32_t prog[P10_RECON_WARPS];     // (bands finished by the warp) * (nblocks+1) + blocks of the current band's last row in memory
    const P10Image& im = imgs[blockIdx.x];
    if (im.channels != CH || !status[im.image]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < P10_RECON_WARPS) prog[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t W = im.w, H = im.h, WP = im.wp;
    const uint32_t nblocks = WP >> 3;
    const unsigned long long ndec = ndecoded[blockIdx.x];
    uint16_t* out16 = (uint16_t*)im.out;
    const bool vec_ok = (W & 7) == 0 && (((uintptr_t)im.out) & 15) == 0;

    const int pw = (warp + P10_RECON_WARPS - 1) % P10_RECON_WARPS;      // the warp that owns the band above mine
    uint32_t kband = 0;
    for (uint32_t y0 = (uint32_t)warp * 32; y0 < H; y0 += 32 * P10_RECON_WARPS, ++kband) {
        const uint32_t y = y0 + lane;
        const bool row_ok = y < H;
        const int lastlane = (int)min(31u, H - 1 - y0);
        const uint32_t kprev_base = (y0 ? (y0 / 32 - 1) / P10_RECON_WARPS : 0) * (nblocks + 1);
        // start block of every lane: one block behind the lane above, later when the row starts inside a run
        const uint32_t info = (row_ok && (unsigned long long)y * W < ndec) ? im.rowinfo[y] : 0u;
        const bool wrap = info != 0;
        const int xn = (int)info - 2;                                   // column of the row above the first pixel copies (-1: any)
        const bool frommem = row_ok && y > 0 && (lane == 0 || wrap);
        uint32_t skew = lane == 0 ? 0u : 1u;
        if (wrap && lane > 0) skew = max((uint32_t)(max(xn, 0) >> 3) + 1u, 2u);
        uint32_t start = skew;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, start, d); if (lane >= d) start += n; }
        const uint32_t total = __shfl_sync(0xffffffffu, start, 31) + nblocks;

        int left = 0;                   // l of the previous pixel in raster order; row 0 starts from the initial predictor 0
        int upleft = 0;
        int prevres[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) prevres[i] = 0;
        const uint32_t* __restrict__ rrow = im.recs + (size_t)y * WP;
        uint4 ra = make_uint4(0, 0, 0, 0), rb = ra;                      // records of the current block
        int upm[8];                                                      // row above from memory (next block, prefetched)
#pragma unroll
        for (int i = 0; i < 8; ++i) upm[i] = 0;

        for (uint32_t T = 0; T < total; ++T) {
            const bool active = row_ok && T >= start && T < start + nblocks;
            const uint32_t blk = T - start, x0 = blk << 3;
            int up[8];
            // the row above: shuffle the previous block results of the lane above, or memory
#pragma unroll
            for (int i = 0; i < 8; ++i) up[i] = __shfl_up_sync(0xffffffffu, prevres[i], 1);
            if (active) {
                if (lane == 0 && y0 > 0) {
                    // the row above belongs to another warp: wait until it is in memory as far as this step reads it
                    uint32_t need = min(blk + 2, nblocks);
                    if (blk == 0 && wrap) need = max(need, min((uint32_t)(max(xn, 0) >> 3) + 1u, nblocks));
                    while (prog[pw] < kprev_base + need) __nanosleep(40);
                }
                if (blk == 0) {
                    ra = __ldg((const uint4*)rrow); rb = __ldg((const uint4*)rrow + 1);
                    if (frommem) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) upm[i] = (uint32_t)i < W ? p10_out_l<CH>(im.out, W, y - 1, i) : 0;
                        if (wrap) left = p10_out_l<CH>(im.out, W, y - 1, (uint32_t)max(xn, 0));
                    }
                }
                if (frommem) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) up[i] = upm[i];
                }
                // prefetch the next block
                uint4 na = ra, nb = rb;
                if (blk + 1 < nblocks) {
                    na = __ldg((const uint4*)(rrow + x0 + 8)); nb = __ldg((const uint4*)(rrow + x0 + 12));
                    if (frommem) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) upm[i] = x0 + 8 + i < W ? p10_out_l<CH>(im.out, W, y - 1, x0 + 8 + i) : 0;
                    }
                }
                const uint32_t recs[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
                uint32_t o[8];
                const unsigned long long gi0 = (unsigned long long)y * W + x0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t r = recs[i];
                    const int val = (int)(r & 1023u);
                    const uint32_t kind = r & (3u << 20);
                    int pred;
                    if (y == 0) pred = left;
                    else if (x0 + i == 0) pred = up[0];
                    else pred = p10_med(left, up[i], i ? up[i - 1] : upleft);
                    int l = (pred + p10_sext((uint32_t)val, 10)) & 1023;
                    if (kind == P10_KIND_COPY) l = left;
                    if (kind == P10_KIND_LIT) l = val;
                    const bool dec = gi0 + i < ndec;
                    if (!dec) l = 0;
                    const uint32_t a = dec ? (r >> 10) & 1023u : 0u;
                    left = l;
                    prevres[i] = l;
                    const uint32_t l16 = (uint32_t)((l << 6) | (l >> 4)), a16 = (a << 6) | (a >> 4);
                    o[i] = CH == 2 ? (l16 | (a16 << 16)) : l16;
                }
                upleft = up[7];
                if (CH == 2) {
                    uint32_t* d = (uint32_t*)(out16 + ((size_t)y * W + x0) * 2);
                    if (vec_ok) { ((uint4*)d)[0] = make_uint4(o[0], o[1], o[2], o[3]); ((uint4*)d)[1] = make_uint4(o[4], o[5], o[6], o[7]); }
                    else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) if (x0 + i < W) d[i] = o[i];
                    }
                } else {
                    uint16_t* d = out16 + (size_t)y * W + x0;
                    if (vec_ok) *(uint4*)d = make_uint4(o[0] | (o[1] << 16), o[2] | (o[3] << 16), o[4] | (o[5] << 16), o[6] | (o[7] << 16));
                    else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) if (x0 + i < W) d[i] = (uint16_t)o[i];
                    }
                }
                ra = na; rb = nb;
                if (lane == lastlane) { __threadfence_block(); prog[warp] = kband * (nblocks + 1) + blk + 1; }
            }
            __syncwarp();       // orders this block's stores before the loads of the lanes that read rows from memory
        }
        __syncwarp();
        if (lane == 0) { __threadfence_block(); prog[warp] = (kband + 1) * (nblocks + 1); }
    }
}
