// lz_resolve.cuh -- deferred LZ77 match resolution shared by the DEFLATE (inflate_par.cuh) and LZ4 (qoix.cu) decoders.
//
// The parallel decoders write literals to their final positions and cannot copy a match on the spot (its source may
// not be written yet), so a match is parked as a small record in the first bytes of the hole it will fill, and its
// start is flagged in a bitmap (1 bit per output byte). One warp per stream then performs the copies in stream order:
// the bitmap is scanned 4096 output bytes (128 words, 4 per lane) at a time; the matches found are taken 32 at a
// time (one per lane); the records of the next 32 matches are loaded before the current 32 are copied, so that only
// the source loads sit on the critical path. Inside a batch the short matches are copied together, one per lane: a
// source byte that is itself produced by a match of the batch is traced back (by address arithmetic over warp
// shuffles) to a byte that is already final, so the batch needs one memory round trip whatever its inner
// dependencies; long matches are copied by the whole warp.
#pragma once
#include <stdint.h>

namespace gb {

enum { LZR_DEFLATE = 0,     // record: len-3 (1 byte), dist-1 (2 bytes LE); len 3..258
       LZR_LZ4 = 1 };       // record: len (2 bytes LE, 4..65535), dist (2 bytes LE)

struct LzMatch { uint32_t dst, len, dist; };

template <int FMT>
__device__ __forceinline__ LzMatch lzr_locate(const uint8_t* out, const uint32_t (&w)[4], uint32_t incl, uint32_t wb,
                                              uint32_t m, bool valid, int lane)
{
    uint32_t lo = 0;
#pragma unroll
    for (int step = 16; step; step >>= 1) {
        const uint32_t v = __shfl_sync(0xffffffffu, incl, (int)(lo + step - 1));
        if (v <= m) lo += step;
    }
    uint32_t excl = __shfl_sync(0xffffffffu, incl, (int)((lo + 31) & 31));
    if (lo == 0) excl = 0;
    const uint32_t x0 = __shfl_sync(0xffffffffu, w[0], (int)lo), x1 = __shfl_sync(0xffffffffu, w[1], (int)lo);
    const uint32_t x2 = __shfl_sync(0xffffffffu, w[2], (int)lo), x3 = __shfl_sync(0xffffffffu, w[3], (int)lo);
    LzMatch M; M.dst = 0; M.len = 0; M.dist = 1;
    if (valid) {
        uint32_t r = m - excl, k = 0, x = x0;
        const uint32_t c0 = __popc(x0), c1 = __popc(x1), c2 = __popc(x2);
        if (r >= c0) { r -= c0; k = 1; x = x1; if (r >= c1) { r -= c1; k = 2; x = x2; if (r >= c2) { r -= c2; k = 3; x = x3; } } }
        const uint32_t bit = __fns(x, 0, (int)(r + 1));
        M.dst = (wb + lo * 4 + k) * 32 + bit;
        if (FMT == LZR_DEFLATE) {
            M.len = (uint32_t)out[M.dst] + 3;
            M.dist = ((uint32_t)out[M.dst + 1] | ((uint32_t)out[M.dst + 2] << 8)) + 1;
        } else {
            M.len = (uint32_t)out[M.dst] | ((uint32_t)out[M.dst + 1] << 8);
            M.dist = (uint32_t)out[M.dst + 2] | ((uint32_t)out[M.dst + 3] << 8);
        }
    }
    return M;
}

__device__ __forceinline__ void lzr_copy_lane(uint8_t* out, uint32_t dst, uint32_t src, uint32_t len, uint32_t dist)
{
    const bool overlap = dist < len;
    for (uint32_t i = 0; i < len; i += 8) {
        uint8_t t[8];
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) {
            const uint32_t x = i + k;
            if (x < len) t[k] = out[src + (overlap ? x % dist : x)];
        }
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) if (i + k < len) out[dst + i + k] = t[k];
    }
}

constexpr uint32_t LZR_SHORT = 16;      // matches up to this length are copied one per lane, longer ones by the whole warp

// One run of consecutive short matches (lanes lo..hi-1 of the batch, ascending destinations). A source byte that lies
// inside the destination of an earlier match of the run (or of the match itself: overlapping copy) is not in memory
// yet -- but its value is known to equal the byte `dist` further back, so the source ADDRESS is chased back through
// the matches of the run (warp shuffles only) until it reaches a byte that is final: before the run's first
// destination, or in a literal gap. All loads of the run are then independent: one memory round trip per run.
__device__ __forceinline__ void lzr_copy_run(uint8_t* out, uint32_t dst, uint32_t len, uint32_t dist, uint32_t runm, int lane)
{
    const bool mine = (runm >> lane) & 1;
    const int lo0 = __ffs(runm) - 1;
    const uint32_t D0 = __shfl_sync(0xffffffffu, dst, lo0);
    // search key, ascending over the lanes: 0 below the run (matches already copied), the destinations, ~0 above
    const uint32_t key = mine ? dst : (lane < lo0 ? 0u : 0xffffffffu);
    const uint32_t maxlen = __reduce_max_sync(0xffffffffu, mine ? len : 0u);
    uint8_t t[LZR_SHORT];
#pragma unroll
    for (uint32_t k = 0; k < LZR_SHORT; ++k) {
        if (k < maxlen) {                                       // warp-uniform
            const bool act = mine && k < len;
            uint32_t a = dst - dist + k;
            bool chasing = act && a >= D0;
            while (__any_sync(0xffffffffu, chasing)) {
                // largest lane i of the run with dst_i <= a
                uint32_t i = 0;
#pragma unroll
                for (int step = 16; step; step >>= 1) {
                    const uint32_t v = __shfl_sync(0xffffffffu, key, (int)((i + step) & 31));
                    if (i + step < 32 && v <= a) i += step;
                }
                const uint32_t di = __shfl_sync(0xffffffffu, key, (int)i), li = __shfl_sync(0xffffffffu, len, (int)i);
                const uint32_t ti = __shfl_sync(0xffffffffu, dist, (int)i);
                if (chasing) {
                    if (((runm >> i) & 1) && di <= a && a - di < li) { a -= ti; chasing = a >= D0; }   // inside match i: same byte, dist_i back
                    else chasing = false;                                           // a literal byte: final
                }
            }
            if (act) t[k] = out[a];
        }
    }
#pragma unroll
    for (uint32_t k = 0; k < LZR_SHORT; ++k) if (mine && k < len) out[dst + k] = t[k];
}

__device__ __forceinline__ void lzr_copy_batch(uint8_t* out, const LzMatch& M, bool valid, int lane)
{
    const uint32_t dst = M.dst, len = M.len, dist = M.dist;
    uint32_t rem = __ballot_sync(0xffffffffu, valid);
    const uint32_t longm = __ballot_sync(0xffffffffu, valid && len > LZR_SHORT);
    while (rem) {
        const int first = __ffs(rem) - 1;
        if ((longm >> first) & 1) {
            // a long match: the whole warp copies it (x % dist: every byte comes from the dist bytes before the
            // destination, so no chunk depends on another)
            const uint32_t d = __shfl_sync(0xffffffffu, dst, first), ln = __shfl_sync(0xffffffffu, len, first);
            const uint32_t di = __shfl_sync(0xffffffffu, dist, first), s = d - di;
            const bool ov = di < ln;
            for (uint32_t c = 0; c < ln; c += 256) {
                uint8_t t[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) { const uint32_t x = c + (uint32_t)k * 32 + lane; if (x < ln) t[k] = out[s + (ov ? x % di : x)]; }
#pragma unroll
                for (int k = 0; k < 8; ++k) { const uint32_t x = c + (uint32_t)k * 32 + lane; if (x < ln) out[d + x] = t[k]; }
            }
            rem &= ~(1u << first);
        } else {
            // the run of short matches up to the next long one
            const uint32_t ahead = longm & rem;
            const uint32_t runm = ahead ? (rem & ((1u << (__ffs(ahead) - 1)) - 1)) : rem;
            lzr_copy_run(out, dst, len, dist, runm, lane);
            rem &= ~runm;
        }
        __syncwarp();
    }
}

// All 32 lanes of one warp. `bm` must be 16-byte aligned and readable up to the next multiple of 128 words.
template <int FMT>
__device__ inline void lz_resolve_stream(uint8_t* out, uint32_t n, const uint32_t* bm, int lane)
{
    const uint32_t nw = (n + 31) >> 5;
    auto load_window = [&](uint32_t wb, uint32_t (&w)[4]) {
        const uint32_t i0 = wb + lane * 4;
        const uint4 v = i0 < nw ? *(const uint4*)(bm + i0) : make_uint4(0, 0, 0, 0);
        w[0] = i0 < nw ? v.x : 0; w[1] = i0 + 1 < nw ? v.y : 0; w[2] = i0 + 2 < nw ? v.z : 0; w[3] = i0 + 3 < nw ? v.w : 0;
    };
    uint32_t w[4], incl = 0, total = 0, wb = 0, m0 = 0;
    bool have = false;
    LzMatch cur; bool curv = false;
    auto next_batch = [&](LzMatch& M, bool& v) -> bool {
        for (;;) {
            if (have && m0 < total) {
                const uint32_t m = m0 + lane;
                v = m < total;
                M = lzr_locate<FMT>(out, w, incl, wb, m, v, lane);
                m0 += 32;
                return true;
            }
            if (have) wb += 128;
            if (wb >= nw) return false;
            load_window(wb, w);
            incl = __popc(w[0]) + __popc(w[1]) + __popc(w[2]) + __popc(w[3]);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
            total = __shfl_sync(0xffffffffu, incl, 31);
            m0 = 0; have = true;
        }
    };
    bool more = next_batch(cur, curv);
    while (more) {
        LzMatch nxt; bool nxtv = false;
        more = next_batch(nxt, nxtv);
        lzr_copy_batch(out, cur, curv, lane);
        cur = nxt; curv = nxtv;
    }
}

} // namespace gb
