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

// ---------------------------------------------------------------------------------------------------------------------
// Chunked resolver (round 2). The stream is walked in 4 KB chunks that live in shared memory while their matches are
// resolved: the chunk (literals in place, match records parked in their holes) is loaded with coalesced vectors, the
// matches that start in it are taken 32 at a time, one per lane, and the chunk is written back once. Inside a batch a
// match may copy as soon as its source ends before the first destination that is still open (everything before that
// point is final): independent matches -- the common case on photographic data, where distances are spread over the
// whole window -- all copy in the first round, chains of dependent matches take one round per link, with the data in
// shared memory instead of a DRAM round trip per link. Sources before the chunk are read from the output buffer
// (final: earlier chunks were written back by this same warp). Matches longer than 32 bytes are copied by the whole
// warp (x % dist: every byte comes from the dist bytes before the destination).
constexpr uint32_t LZC_BYTES = 4096, LZC_EXTRA = 320, LZC_SPAN = LZC_BYTES + LZC_EXTRA, LZC_BUF = LZC_SPAN + 16;

template <int FMT>
__device__ inline void lz_resolve_stream_chunked(uint8_t* out, uint32_t n, const uint32_t* bm, int lane, uint8_t* buf0 /* LZC_BUF bytes, 16-aligned */)
{
    const uint32_t nw = (n + 31) >> 5;
    for (uint32_t c0 = 0; c0 < n; c0 += LZC_BYTES) {
        const uint32_t cend = min(c0 + LZC_SPAN, n);                // bytes of the stream held in the buffer: [c0, cend)
        // the buffer starts at the 16-byte boundary at or below out + c0, so that global accesses are aligned vectors
        // whatever the alignment of `out`; the up to 15 bytes before c0 are final bytes of the previous chunk (or, for
        // c0 = 0, bytes of the allocation before `out`: device allocations are 256-byte aligned)
        const uint32_t mis = (uint32_t)((uintptr_t)(out + c0) & 15);
        uint8_t* const buf = buf0 + mis;                             // buf[x - c0] for stream position x >= c0 - mis
        uint8_t* const gbase = out + c0 - mis;                       // 16-byte aligned
        const uint32_t span = cend - c0 + mis;
        // ---- bitmap window of the chunk: 128 words, 4 per lane
        uint32_t w[4];
        {
            const uint32_t i0 = (c0 >> 5) + lane * 4;
            const uint4 v = i0 < nw ? *(const uint4*)(bm + i0) : make_uint4(0, 0, 0, 0);
            w[0] = i0 < nw ? v.x : 0; w[1] = i0 + 1 < nw ? v.y : 0; w[2] = i0 + 2 < nw ? v.z : 0; w[3] = i0 + 3 < nw ? v.w : 0;
        }
        uint32_t incl = __popc(w[0]) + __popc(w[1]) + __popc(w[2]) + __popc(w[3]);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;                                   // nothing starts here: the chunk is final as it stands
        // ---- load
        for (uint32_t i = lane * 16; i < span; i += 512) *(uint4*)(buf0 + i) = __ldcg((const uint4*)(gbase + i));   // reads <= 15 bytes past n: slack
        __syncwarp();
        for (uint32_t m0 = 0; m0 < total; m0 += 32) {
            const uint32_t m = m0 + lane;
            const bool valid = m < total;
            // locate match m of the window (same search as lzr_locate) and read its record from the buffer
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
            uint32_t dst = 0xffffffffu, len = 0, dist = 1;
            if (valid) {
                uint32_t r = m - excl, k = 0, x = x0;
                const uint32_t p0 = __popc(x0), p1 = __popc(x1), p2 = __popc(x2);
                if (r >= p0) { r -= p0; k = 1; x = x1; if (r >= p1) { r -= p1; k = 2; x = x2; if (r >= p2) { r -= p2; k = 3; x = x3; } } }
                const uint32_t bit = __fns(x, 0, (int)(r + 1));
                dst = c0 + (lo * 4 + k) * 32 + bit;
                const uint8_t* rec = buf + (dst - c0);
                if (FMT == LZR_DEFLATE) { len = (uint32_t)rec[0] + 3; dist = ((uint32_t)rec[1] | ((uint32_t)rec[2] << 8)) + 1; }
                else { len = (uint32_t)rec[0] | ((uint32_t)rec[1] << 8); dist = (uint32_t)rec[2] | ((uint32_t)rec[3] << 8); }
            }
            // a record that would read before the stream or write past it cannot come from an accepted stream: skip it
            bool done = !valid || dist == 0 || dist > dst || dst + len > n;
            __syncwarp();
            for (;;) {
                const uint32_t F = __reduce_min_sync(0xffffffffu, done ? 0xffffffffu : dst);
                if (F == 0xffffffffu) break;
                const int fl = __ffs(__ballot_sync(0xffffffffu, !done && dst == F)) - 1;
                const uint32_t flen = __shfl_sync(0xffffffffu, len, fl);
                if (flen > 32) {
                    // the first open match is long: the whole warp copies it
                    const uint32_t d = F, di = __shfl_sync(0xffffffffu, dist, fl), s = d - di;
                    const bool ov = di < flen;
                    if (d + flen <= c0 + LZC_SPAN) {
                        for (uint32_t x = lane; x < flen; x += 32) {
                            const uint32_t a = s + (ov ? x % di : x);
                            buf[d - c0 + x] = a >= c0 ? buf[a - c0] : __ldcg(out + a);
                        }
                    } else {
                        // longer than the buffer's slack (LZ4 only): through the output buffer. Everything resolved so
                        // far goes back first, the part of the destination that the buffer covers is reloaded after.
                        __syncwarp();
                        for (uint32_t i = lane; i < cend - c0; i += 32) out[c0 + i] = buf[i];
                        __threadfence_block();
                        __syncwarp();
                        for (uint32_t c = 0; c < flen; c += 256) {
                            uint8_t t[8];
#pragma unroll
                            for (int q = 0; q < 8; ++q) { const uint32_t x = c + (uint32_t)q * 32 + lane; if (x < flen) t[q] = __ldcg(out + s + (ov ? x % di : x)); }
#pragma unroll
                            for (int q = 0; q < 8; ++q) { const uint32_t x = c + (uint32_t)q * 32 + lane; if (x < flen) out[d + x] = t[q]; }
                        }
                        __threadfence_block();
                        __syncwarp();
                        for (uint32_t i = d - c0 + lane; i < cend - c0; i += 32) buf[i] = __ldcg(out + c0 + i);
                    }
                    if (lane == fl) done = true;
                    __syncwarp();
                    continue;
                }
                // short matches whose source ends before the first open destination copy now, one per lane
                const uint32_t s = dst - dist;
                const bool ready = !done && len <= 32 && (s + min(len, dist) <= F || dst == F);
                if (ready) {
                    uint8_t* d8 = buf + (dst - c0);
                    if (s + len <= c0) {
                        // source entirely before the chunk: independent loads from the output buffer
                        uint8_t t[32];
#pragma unroll
                        for (uint32_t k = 0; k < 32; ++k) if (k < len) t[k] = __ldcg(out + s + k);
#pragma unroll
                        for (uint32_t k = 0; k < 32; ++k) if (k < len) d8[k] = t[k];
                    } else {
                        // byte by byte in stream order (an overlapping copy reads what it has just written)
                        for (uint32_t k = 0; k < len; ++k) { const uint32_t a = s + k; d8[k] = a >= c0 ? buf[a - c0] : __ldcg(out + a); }
                    }
                    done = true;
                }
                __syncwarp();
            }
        }
        // ---- write back [c0, cend): resolved bytes, untouched literals, and the still-parked records of the next chunk
        __syncwarp();
        {
            const uint32_t nfull = span & ~15u;
            for (uint32_t i = lane * 16; i < nfull; i += 512) *(uint4*)(gbase + i) = *(const uint4*)(buf0 + i);
            if (nfull + lane < span) gbase[nfull + lane] = buf0[nfull + lane];       // at most 15 tail bytes, never past n
        }
        __threadfence_block();
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Chunked resolver, second form (round 2, PNG). Same 4 KB chunks in shared memory, but the work of a chunk is ordered
// by what it waits for instead of by stream position:
//   1. enumerate  every lane walks the set bits of its 4 bitmap words and writes (offset, length, distance) of its
//                 matches to a list in shared memory (the lane's slot range comes from the popcount scan);
//   2. far        matches whose whole source lies before the chunk (three quarters of them on photographic PNG data:
//                 the row above is 7681 bytes back) depend on nothing that is still open: all of them are copied at
//                 once, four per lane in flight, each as two or three aligned 8-byte loads realigned with a shift --
//                 one L2 round trip per 128 matches instead of one per dependency round of 32;
//   3. near       the remaining matches (source inside the chunk, or longer than 16 bytes) are taken in stream order,
//                 32 at a time, exactly like lz_resolve_stream_chunked: a match copies as soon as its source ends
//                 before the first destination that is still open; all the data involved is in shared memory.
constexpr uint32_t LZ3_MAXM = LZC_BYTES / 3 + 3;
constexpr uint32_t LZ3_FAR_LEN = 16;
constexpr int LZ3_G = 4;

struct Lz3Shared {
    __align__(16) uint8_t buf[LZC_BUF];
    uint2 list[LZ3_MAXM];                   // x: offset in the chunk | length << 12 (0: skip), y: distance
    uint16_t near_idx[LZ3_MAXM + 1];
};

template <int FMT>
__device__ inline void lz_resolve_stream_v3(uint8_t* out, uint32_t n, const uint32_t* bm, int lane, Lz3Shared& S)
{
    const uint32_t nw = (n + 31) >> 5;
    for (uint32_t c0 = 0; c0 < n; c0 += LZC_BYTES) {
        const uint32_t cend = min(c0 + LZC_SPAN, n);
        const uint32_t mis = (uint32_t)((uintptr_t)(out + c0) & 15);
        uint8_t* const buf = S.buf + mis;                            // buf[x - c0] for stream position x >= c0 - mis
        uint8_t* const gbase = out + c0 - mis;                       // 16-byte aligned
        const uint32_t span = cend - c0 + mis;
        uint32_t w[4];
        {
            const uint32_t i0 = (c0 >> 5) + lane * 4;
            const uint4 v = i0 < nw ? *(const uint4*)(bm + i0) : make_uint4(0, 0, 0, 0);
            w[0] = i0 < nw ? v.x : 0; w[1] = i0 + 1 < nw ? v.y : 0; w[2] = i0 + 2 < nw ? v.z : 0; w[3] = i0 + 3 < nw ? v.w : 0;
        }
        const uint32_t cnt = __popc(w[0]) + __popc(w[1]) + __popc(w[2]) + __popc(w[3]);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
        const uint32_t total = min(__shfl_sync(0xffffffffu, incl, 31), LZ3_MAXM);
        if (total == 0) continue;                                   // nothing starts here: the chunk is final as it stands
        for (uint32_t i = lane * 16; i < span; i += 512) *(uint4*)(S.buf + i) = __ldcg((const uint4*)(gbase + i));
        __syncwarp();
        // ---- 1. enumerate
        {
            uint32_t idx = incl - cnt;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t x = w[k];
                const uint32_t base = ((uint32_t)lane * 4 + (uint32_t)k) * 32;
                while (x) {
                    const uint32_t p = base + (uint32_t)__ffs(x) - 1;
                    x &= x - 1;
                    const uint8_t* rec = buf + p;
                    uint32_t len, dist;
                    if (FMT == LZR_DEFLATE) { len = (uint32_t)rec[0] + 3; dist = ((uint32_t)rec[1] | ((uint32_t)rec[2] << 8)) + 1; }
                    else { len = (uint32_t)rec[0] | ((uint32_t)rec[1] << 8); dist = (uint32_t)rec[2] | ((uint32_t)rec[3] << 8); }
                    const uint32_t dst = c0 + p;
                    // a record that would read before the stream or write past it cannot come from an accepted stream: skip it
                    if (dist == 0 || dist > dst || dst + len > n) len = 0;
                    if (idx < LZ3_MAXM) S.list[idx] = make_uint2(p | (len << 12), dist);
                    ++idx;
                }
            }
        }
        __syncwarp();
        // ---- 2. far matches, LZ3_G per lane in flight; the others are queued in stream order for step 3
        uint32_t nnear = 0;
        for (uint32_t m0 = 0; m0 < total; m0 += 32 * LZ3_G) {
            uint32_t p[LZ3_G], len[LZ3_G];
            uint64_t q0[LZ3_G], q1[LZ3_G], q2[LZ3_G];
            bool far[LZ3_G], nearf[LZ3_G];
            uint32_t sh[LZ3_G];
#pragma unroll
            for (int g = 0; g < LZ3_G; ++g) {
                const uint32_t m = m0 + (uint32_t)g * 32 + (uint32_t)lane;
                const uint2 e = m < total ? S.list[m] : make_uint2(0, 1);
                p[g] = e.x & 4095; len[g] = e.x >> 12;
                const uint32_t s = c0 + p[g] - e.y;
                far[g] = len[g] != 0 && len[g] <= LZ3_FAR_LEN && s + len[g] <= c0;
                nearf[g] = len[g] != 0 && !far[g];
                q0[g] = q1[g] = q2[g] = 0; sh[g] = 0;
                if (far[g]) {
                    const uintptr_t a = (uintptr_t)(out + s);
                    const uint64_t* a8 = (const uint64_t*)(a & ~(uintptr_t)7);
                    const uint32_t o = (uint32_t)(a & 7);
                    sh[g] = o * 8;
                    q0[g] = __ldcg(a8);
                    if (o + len[g] > 8) q1[g] = __ldcg(a8 + 1);
                    if (o + len[g] > 16) q2[g] = __ldcg(a8 + 2);
                }
            }
#pragma unroll
            for (int g = 0; g < LZ3_G; ++g) {
                if (far[g]) {
                    uint64_t lo = q0[g] >> sh[g], hi = q1[g] >> sh[g];
                    if (sh[g]) { lo |= q1[g] << (64 - sh[g]); hi |= q2[g] << (64 - sh[g]); }
                    uint8_t* d8 = buf + p[g];
#pragma unroll
                    for (uint32_t k = 0; k < 8; ++k) if (k < len[g]) d8[k] = (uint8_t)(lo >> (8 * k));
                    if (len[g] > 8) {
#pragma unroll
                        for (uint32_t k = 0; k < 8; ++k) if (8 + k < len[g]) d8[8 + k] = (uint8_t)(hi >> (8 * k));
                    }
                }
                const uint32_t nm = __ballot_sync(0xffffffffu, nearf[g]);
                if (nearf[g]) S.near_idx[nnear + __popc(nm & ((1u << lane) - 1))] = (uint16_t)(m0 + (uint32_t)g * 32 + (uint32_t)lane);
                nnear += __popc(nm);
            }
        }
        __syncwarp();
        // ---- 3. near matches in stream order
        for (uint32_t b0 = 0; b0 < nnear; b0 += 32) {
            const bool valid = b0 + lane < nnear;
            const uint2 e = valid ? S.list[S.near_idx[b0 + lane]] : make_uint2(0, 1);
            const uint32_t dst = valid ? c0 + (e.x & 4095) : 0xffffffffu, len = e.x >> 12, dist = e.y;
            bool done = !valid;
            for (;;) {
                const uint32_t F = __reduce_min_sync(0xffffffffu, done ? 0xffffffffu : dst);
                if (F == 0xffffffffu) break;
                const int fl = __ffs(__ballot_sync(0xffffffffu, !done && dst == F)) - 1;
                const uint32_t flen = __shfl_sync(0xffffffffu, len, fl);
                if (flen > 32) {
                    // the first open match is long: the whole warp copies it
                    const uint32_t d = F, di = __shfl_sync(0xffffffffu, dist, fl), s = d - di;
                    const bool ov = di < flen;
                    if (d + flen <= c0 + LZC_SPAN) {
                        if (!ov) {
                            for (uint32_t x = lane; x < flen; x += 32) { const uint32_t a = s + x; buf[d - c0 + x] = a >= c0 ? buf[a - c0] : __ldcg(out + a); }
                        } else {
                            // periodic: byte x equals source byte x mod di; all sources lie before d (final)
                            for (uint32_t x = lane; x < flen; x += 32) { const uint32_t a = s + x % di; buf[d - c0 + x] = a >= c0 ? buf[a - c0] : __ldcg(out + a); }
                        }
                    } else {
                        // longer than the buffer's slack (LZ4 only): through the output buffer. Everything resolved so
                        // far goes back first, the part of the destination that the buffer covers is reloaded after.
                        __syncwarp();
                        for (uint32_t i = lane; i < cend - c0; i += 32) out[c0 + i] = buf[i];
                        __threadfence_block();
                        __syncwarp();
                        for (uint32_t c = 0; c < flen; c += 256) {
                            uint8_t t[8];
#pragma unroll
                            for (int q = 0; q < 8; ++q) { const uint32_t x = c + (uint32_t)q * 32 + lane; if (x < flen) t[q] = __ldcg(out + s + (ov ? x % di : x)); }
#pragma unroll
                            for (int q = 0; q < 8; ++q) { const uint32_t x = c + (uint32_t)q * 32 + lane; if (x < flen) out[d + x] = t[q]; }
                        }
                        __threadfence_block();
                        __syncwarp();
                        for (uint32_t i = d - c0 + lane; i < cend - c0; i += 32) buf[i] = __ldcg(out + c0 + i);
                    }
                    if (lane == fl) done = true;
                    __syncwarp();
                    continue;
                }
                // short matches whose source ends before the first open destination copy now, one per lane
                const uint32_t s = dst - dist;
                const bool ready = !done && len <= 32 && (s + min(len, dist) <= F || dst == F);
                if (ready) {
                    uint8_t* d8 = buf + (dst - c0);
                    if (dist >= len) {
                        // no self-overlap: independent loads, 8 at a time
                        for (uint32_t k0 = 0; k0 < len; k0 += 8) {
                            uint8_t t[8];
#pragma unroll
                            for (uint32_t k = 0; k < 8; ++k) if (k0 + k < len) { const uint32_t a = s + k0 + k; t[k] = a >= c0 ? buf[a - c0] : __ldcg(out + a); }
#pragma unroll
                            for (uint32_t k = 0; k < 8; ++k) if (k0 + k < len) d8[k0 + k] = t[k];
                        }
                    } else {
                        // overlapping copy: the first `dist` bytes come from the source, the rest repeats them
                        uint8_t t[8];
                        for (uint32_t k0 = 0; k0 < dist; k0 += 8) {
#pragma unroll
                            for (uint32_t k = 0; k < 8; ++k) if (k0 + k < dist) { const uint32_t a = s + k0 + k; t[k] = a >= c0 ? buf[a - c0] : __ldcg(out + a); }
#pragma unroll
                            for (uint32_t k = 0; k < 8; ++k) if (k0 + k < dist) d8[k0 + k] = t[k];
                        }
                        for (uint32_t k = dist; k < len; ++k) d8[k] = d8[k - dist];
                    }
                    done = true;
                }
                __syncwarp();
            }
        }
        // ---- write back [c0, cend): resolved bytes, untouched literals, and the still-parked records of the next chunk
        __syncwarp();
        {
            const uint32_t nfull = span & ~15u;
            for (uint32_t i = lane * 16; i < nfull; i += 512) *(uint4*)(gbase + i) = *(const uint4*)(S.buf + i);
            if (nfull + lane < span) gbase[nfull + lane] = S.buf[nfull + lane];       // at most 15 tail bytes, never past n
        }
        __threadfence_block();
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Chunked resolver, CTA form with pointer jumping (round 2, PNG; matches of at most LZC_EXTRA bytes, i.e. DEFLATE).
// Profiles of the one-warp chunked resolver on the adaptive-filter PNG workload (profiles/r2_resolve_*): a lone warp runs
// 13.6 k warp instructions per 4 KB chunk at one instruction every ~10 cycles (143 ms per stream whatever the batch
// size), 60 % of them in ordered rounds with 3.4 of 32 lanes active -- the matches of filtered image rows form long
// dependency chains (a match copies what the match before it has just produced; distances of 4-8 bytes), and every
// link of a chain was a round. A first CTA version that chased source addresses back through the open matches (one
// hop per link) removed the ordering but paid one hop per link and byte: 15.7 k instructions per chunk, 90 ms.
// This version resolves the chains in O(log depth) uniform passes instead:
//   * NT threads per stream; the chunk and a pointer per byte (16 bits, chunk-relative) sit in shared memory; a byte
//     that is final points to itself;
//   * far matches (whole source before the chunk: final data) are copied at once, with aligned 8-byte loads (<= 32
//     bytes, one match per thread, two in flight) or by the whole warp (longer ones, two in flight);
//   * every byte x of any other match gets ptr[x] = x - dist (a source byte before the chunk is fetched on the spot
//     and the byte becomes final); for an overlapping match this chains through the match itself;
//   * pointer jumping, ptr[x] = ptr[ptr[x]], in place, until nothing changes: every pass at least doubles the distance
//     a pointer has travelled, so <= 13 passes whatever the input and 3-5 on image data (in-place updates in ascending
//     order compress most of a chain in one pass); all lanes busy, no divergence, one barrier per pass;
//   * one pass copies buf[x] = buf[ptr[x]]: sources are roots (final bytes), destinations are not -- no hazards.
constexpr int LZ4C_NT = 128;
constexpr uint32_t LZ4C_PTRS = (LZC_SPAN + 32 + 7) & ~7u;

struct Lz4cShared {
    __align__(16) uint8_t buf[LZC_BUF];
    __align__(16) uint16_t ptr[LZ4C_PTRS];  // chunk-relative position of the byte this byte is a copy of (itself: final)
    uint2 list[LZ3_MAXM];                   // x: offset in the chunk | length << 12 (0: skip), y: distance
    uint32_t wsum[LZ4C_NT / 32];
};

template <int FMT>
__device__ inline void lz_resolve_stream_jump(uint8_t* out, uint32_t n, const uint32_t* bm, Lz4cShared& S)
{
    constexpr int NT = LZ4C_NT, NWARP = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nw = (n + 31) >> 5;
    for (uint32_t c0 = 0; c0 < n; c0 += LZC_BYTES) {
        const uint32_t cend = min(c0 + LZC_SPAN, n);
        const uint32_t mis = (uint32_t)((uintptr_t)(out + c0) & 15);
        uint8_t* const buf = S.buf + mis;                            // buf[x - c0] for stream position x >= c0 - mis
        uint8_t* const gbase = out + c0 - mis;                       // 16-byte aligned
        const uint32_t span = cend - c0 + mis;
        __syncthreads();                                             // the previous chunk's write-back has read S.buf
        // ---- match starts: one bitmap word per thread (NT * 32 = LZC_BYTES)
        uint32_t w;
        {
            const uint32_t i = (c0 >> 5) + (uint32_t)tid;
            w = i < nw ? bm[i] : 0u;
            const uint32_t lim = n - min(n, c0 + (uint32_t)tid * 32);             // bits at or beyond n are not matches
            if (lim < 32) w &= lim ? (0xffffffffu >> (32 - lim)) : 0u;
        }
        const uint32_t cnt = __popc(w);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
        if (lane == 31) S.wsum[warp] = incl;
        for (uint32_t i = tid * 16; i < span; i += NT * 16) *(uint4*)(S.buf + i) = __ldcg((const uint4*)(gbase + i));
        // every byte final: ptr[x] = x
        for (uint32_t i = tid * 8; i < LZ4C_PTRS; i += NT * 8) {
            const uint32_t a = i | ((i + 1) << 16);
            *(uint4*)(S.ptr + i) = make_uint4(a, a + 0x00020002u, a + 0x00040004u, a + 0x00060006u);
        }
        __syncthreads();
        uint32_t base = 0, total = 0;
#pragma unroll
        for (int q = 0; q < NWARP; ++q) { const uint32_t t = S.wsum[q]; base += q < warp ? t : 0u; total += t; }
        if (total == 0) continue;                                    // nothing starts here: the chunk is final as it stands
        total = min(total, LZ3_MAXM);
        // ---- 1. enumerate
        {
            uint32_t idx = base + incl - cnt;
            uint32_t x = w;
            while (x) {
                const uint32_t p = (uint32_t)tid * 32 + (uint32_t)__ffs(x) - 1;
                x &= x - 1;
                const uint8_t* rec = buf + p;
                uint32_t len, dist;
                if (FMT == LZR_DEFLATE) { len = (uint32_t)rec[0] + 3; dist = ((uint32_t)rec[1] | ((uint32_t)rec[2] << 8)) + 1; }
                else { len = (uint32_t)rec[0] | ((uint32_t)rec[1] << 8); dist = (uint32_t)rec[2] | ((uint32_t)rec[3] << 8); }
                const uint32_t dst = c0 + p;
                // a record that would read before the stream or write past it (or past the buffer's slack) cannot come
                // from an accepted stream: skip it
                if (dist == 0 || dist > dst || dst + len > n || len > LZC_EXTRA) len = 0;
                if (idx < LZ3_MAXM) S.list[idx] = make_uint2(p | (len << 12), dist);
                ++idx;
            }
        }
        __syncthreads();
        // ---- 2. far matches copy; the bytes of the others point dist back
        bool any_open = false;
        for (uint32_t m0 = 0; m0 < total; m0 += NT * 2) {
            uint32_t p[2], len[2], sh[2], srcpos[2];
            uint64_t q[2][5];
            bool far[2];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const uint32_t m = m0 + (uint32_t)g * NT + (uint32_t)tid;
                const uint2 e = m < total ? S.list[m] : make_uint2(0, 1);
                p[g] = e.x & 4095; len[g] = e.x >> 12;
                const uint32_t s = c0 + p[g] - e.y;
                srcpos[g] = s;
                far[g] = len[g] != 0 && s + len[g] <= c0;
#pragma unroll
                for (int i = 0; i < 5; ++i) q[g][i] = 0;
                sh[g] = 0;
                if (far[g] && len[g] <= 32) {
                    const uintptr_t a = (uintptr_t)(out + s);
                    const uint64_t* a8 = (const uint64_t*)(a & ~(uintptr_t)7);
                    const uint32_t o = (uint32_t)(a & 7);
                    sh[g] = o * 8;
                    q[g][0] = __ldcg(a8);
#pragma unroll
                    for (int i = 1; i < 5; ++i) if (o + len[g] > 8u * i) q[g][i] = __ldcg(a8 + i);
                }
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const bool open = len[g] != 0 && !far[g];
                any_open |= open;
                if (open && len[g] <= 32) {
                    // x - dist, or the byte itself when its source lies before the chunk (a straddling match)
                    const uint32_t back = p[g] + c0 - srcpos[g];         // = dist
                    for (uint32_t k = 0; k < len[g]; ++k) {
                        const uint32_t x = p[g] + k;
                        if (x >= back) S.ptr[x] = (uint16_t)(x - back);
                        else buf[x] = __ldcg(out + srcpos[g] + k);
                    }
                }
                uint32_t lm = __ballot_sync(0xffffffffu, open && len[g] > 32);
                while (lm) {
                    const int l = __ffs(lm) - 1; lm &= lm - 1;
                    const uint32_t pp = __shfl_sync(0xffffffffu, p[g], l), ll = __shfl_sync(0xffffffffu, len[g], l);
                    const uint32_t ss = __shfl_sync(0xffffffffu, srcpos[g], l);
                    const uint32_t back = pp + c0 - ss;
                    for (uint32_t k = lane; k < ll; k += 32) {
                        const uint32_t x = pp + k;
                        if (x >= back) S.ptr[x] = (uint16_t)(x - back);
                        else buf[x] = __ldcg(out + ss + k);
                    }
                }
                if (far[g] && len[g] <= 32) {
                    uint8_t* d8 = buf + p[g];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if ((uint32_t)i * 8 < len[g]) {
                            uint64_t v = q[g][i] >> sh[g];
                            if (sh[g]) v |= q[g][i + 1] << (64 - sh[g]);
#pragma unroll
                            for (uint32_t k = 0; k < 8; ++k) if ((uint32_t)i * 8 + k < len[g]) d8[i * 8 + k] = (uint8_t)(v >> (8 * k));
                        }
                    }
                }
                // long far matches of this warp, two at a time
                uint32_t fm = __ballot_sync(0xffffffffu, far[g] && len[g] > 32);
                while (fm) {
                    const int l0 = __ffs(fm) - 1; fm &= fm - 1;
                    const int l1 = fm ? __ffs(fm) - 1 : l0;
                    const bool two = fm != 0; fm &= fm - 1;
                    const uint32_t pa = __shfl_sync(0xffffffffu, p[g], l0), la = __shfl_sync(0xffffffffu, len[g], l0), sa = __shfl_sync(0xffffffffu, srcpos[g], l0);
                    const uint32_t pb = __shfl_sync(0xffffffffu, p[g], l1), lb = two ? __shfl_sync(0xffffffffu, len[g], l1) : 0u, sb = __shfl_sync(0xffffffffu, srcpos[g], l1);
                    uint8_t ta[LZC_EXTRA / 32], tb[LZC_EXTRA / 32];
#pragma unroll
                    for (uint32_t i = 0; i < LZC_EXTRA / 32; ++i) {
                        const uint32_t x = i * 32 + (uint32_t)lane;
                        if (x < la) ta[i] = __ldcg(out + sa + x);
                        if (x < lb) tb[i] = __ldcg(out + sb + x);
                    }
#pragma unroll
                    for (uint32_t i = 0; i < LZC_EXTRA / 32; ++i) {
                        const uint32_t x = i * 32 + (uint32_t)lane;
                        if (x < la) buf[pa + x] = ta[i];
                        if (x < lb) buf[pb + x] = tb[i];
                    }
                }
            }
        }
        if (__syncthreads_or(any_open ? 1 : 0)) {
            // ---- 3. pointer jumping over the chunk, two positions per 32-bit word, in place. A thread owns the pairs tid,
            // tid + NT, ... and keeps one bit per pair: set while one of the two bytes does not point at a root yet. A
            // pair whose two targets are roots is done for good (roots never change), so the passes get shorter.
            const uint32_t npair = (min(cend - c0, LZC_SPAN) + 1) >> 1;
            uint32_t* const ptr32 = (uint32_t*)S.ptr;
            uint32_t act = 0;
            {
                uint32_t it = 0;
                for (uint32_t i = tid; i < npair; i += NT, ++it)
                    if (ptr32[i] != ((2 * i) | ((2 * i + 1) << 16))) act |= 1u << it;
            }
            const uint32_t open0 = act;
            for (int pass = 0; pass < 16; ++pass) {
                bool changed = false;
                uint32_t a = act;
                while (a) {
                    const uint32_t it = (uint32_t)__ffs(a) - 1;
                    a &= a - 1;
                    const uint32_t i = (uint32_t)tid + it * NT;
                    const uint32_t v = ptr32[i];
                    const uint32_t p0 = v & 0xffffu, p1 = v >> 16;
                    const uint32_t q0 = S.ptr[p0], q1 = S.ptr[p1];
                    if (q0 != p0 || q1 != p1) { ptr32[i] = q0 | (q1 << 16); changed = true; }
                    else act &= ~(1u << it);
                }
                if (!__syncthreads_or(changed ? 1 : 0)) break;
            }
            // ---- 4. copy: every open byte from its root (a final byte of an open pair is rewritten with itself)
            {
                uint32_t a = open0;
                while (a) {
                    const uint32_t it = (uint32_t)__ffs(a) - 1;
                    a &= a - 1;
                    const uint32_t i = (uint32_t)tid + it * NT;
                    const uint32_t v = ptr32[i];
                    const uint8_t b0 = buf[v & 0xffffu], b1 = buf[v >> 16];
                    buf[2 * i] = b0; buf[2 * i + 1] = b1;
                }
            }
            __syncthreads();
        }
        // ---- write back [c0, cend): resolved bytes, untouched literals, and the still-parked records of the next chunk
        {
            const uint32_t nfull = span & ~15u;
            for (uint32_t i = tid * 16; i < nfull; i += NT * 16) *(uint4*)(gbase + i) = *(const uint4*)(S.buf + i);
            if (tid < 16 && nfull + tid < span) gbase[nfull + tid] = S.buf[nfull + tid];       // at most 15 tail bytes, never past n
        }
        __threadfence_block();
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------------
// Measured and not kept (round 2): the same resolver with the chunk traffic on the TMA -- cp.async.bulk + mbarrier
// prefetch of chunk c+1 into a second buffer while chunk c jumps pointers, one cp.async.bulk.global.shared write-back
// per chunk, far sources of the previous chunk served from shared memory. 49.0 ms against 45.0 ms per 1024 streams:
// the kernel is bound by instruction issue (57 % issue-active with all 1024 CTAs resident, profiles/r2_resolve_*), not
// by the latency of the chunk load, so hiding that latency buys nothing and the second buffer costs a resident CTA
// per SM.

// ---------------------------------------------------------------------------------------------------------------------
// CTA-wide resolver: one thread per match, exact dependencies. Used where a batch has FEW streams (QOIX: one LZ4 block
// per image, 256 images per GPU), so that a stream gets 16 warps instead of one.
// The chunk sits in shared memory with a finality bit per byte (literals final, match destinations open); the matches
// are taken LZB_THREADS at a time, one per thread; every open match checks the bits of its source range and copies in
// the round in which they are all set, then sets its own. The number of rounds is the depth of the longest dependency
// chain inside the batch, each round costs one barrier; the copies of a round are independent by construction.
// Measured (round 2): LZ4 of 256 x 2048^2 QOIX images 16.4 ms (round-1 resolver) -> 4.3 ms. PNG, 1024 x 1080p with
// 1.1 M matches per image (mean length 6.4, three quarters of the sources more than 4 KB back): round-1 resolver 249
// ms, lz_resolve_stream_chunked (one warp per stream, all streams resident) 112 ms, this one 135-139 ms (it executes
// ~20 k warp instructions per chunk -- every warp runs every round -- and becomes throughput-bound once all 1024
// streams are resident), warps spinning on the bits without barriers 189 ms, a lane-owned-region variant with a
// 16 KB window in shared memory 300 ms (a lone warp issues one dependent instruction every 6-10 cycles). PNG
// therefore uses lz_resolve_stream_chunked.
constexpr int LZB_THREADS = 512;
constexpr uint32_t LZB_FINW = (LZC_SPAN + 31) / 32 + 1;

struct LzbShared {
    __align__(16) uint8_t buf[LZC_BUF];
    uint32_t start[128];                    // match-start bits of the chunk
    uint32_t pref[129];                     // exclusive prefix of their popcounts
    uint32_t wsum[4];
    uint32_t fin[LZB_FINW];                 // finality bits, relative to c0
};

__device__ __forceinline__ bool lzb_range_final(const volatile uint32_t* fin, uint32_t a, uint32_t b)      // bits [a, b), a < b
{
    const uint32_t w0 = a >> 5, w1 = (b - 1) >> 5;
    for (uint32_t w = w0; w <= w1; ++w) {
        uint32_t mask = 0xffffffffu;
        if (w == w0) mask &= 0xffffffffu << (a & 31);
        if (w == w1) mask &= 0xffffffffu >> (31 - ((b - 1) & 31));
        if ((fin[w] & mask) != mask) return false;
    }
    return true;
}
template <bool SET>
__device__ __forceinline__ void lzb_range_mark(uint32_t* fin, uint32_t a, uint32_t b)
{
    const uint32_t w0 = a >> 5, w1 = (b - 1) >> 5;
    for (uint32_t w = w0; w <= w1; ++w) {
        uint32_t mask = 0xffffffffu;
        if (w == w0) mask &= 0xffffffffu << (a & 31);
        if (w == w1) mask &= 0xffffffffu >> (31 - ((b - 1) & 31));
        if (SET) atomicOr(fin + w, mask); else atomicAnd(fin + w, ~mask);
    }
}

// All LZB_THREADS threads of one CTA. `bm` must be 16-byte aligned and readable up to the next multiple of 128 words.
template <int FMT>
__device__ inline void lz_resolve_stream_cta(uint8_t* out, uint32_t n, const uint32_t* bm, LzbShared& S)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t nw = (n + 31) >> 5;
    for (uint32_t c0 = 0; c0 < n; c0 += LZC_BYTES) {
        const uint32_t cend = min(c0 + LZC_SPAN, n);
        const uint32_t mis = (uint32_t)((uintptr_t)(out + c0) & 15);
        uint8_t* const buf = S.buf + mis;                            // buf[x - c0] for stream position x >= c0 - mis
        uint8_t* const gbase = out + c0 - mis;                       // 16-byte aligned
        const uint32_t span = cend - c0 + mis;
        // ---- match starts of the chunk and the prefix of their counts
        __syncthreads();
        if (tid < 128) {
            const uint32_t i = (c0 >> 5) + tid;
            uint32_t w = i < nw ? bm[i] : 0;
            const uint32_t lim = n - min(n, c0 + (uint32_t)tid * 32);             // bits at or beyond n are not matches
            if (lim < 32) w &= lim ? (0xffffffffu >> (32 - lim)) : 0u;
            S.start[tid] = w;
            uint32_t inc = __popc(w);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
            if (lane == 31) S.wsum[tid >> 5] = inc;
            S.pref[tid + 1] = inc;                                    // warp-local inclusive; the warp bases are added below
        }
        __syncthreads();
        if (tid < 128) {
            uint32_t base = 0;
            for (int w4 = 0; w4 < (tid >> 5); ++w4) base += S.wsum[w4];
            S.pref[tid + 1] += base;
            if (tid == 0) S.pref[0] = 0;
        }
        __syncthreads();
        const uint32_t total = S.pref[128];
        if (total == 0) continue;                                     // nothing starts here: the chunk is final as it stands
        // ---- load the chunk, all bytes final until a match claims them
        for (uint32_t i = tid * 16; i < span; i += LZB_THREADS * 16) *(uint4*)(S.buf + i) = __ldcg((const uint4*)(gbase + i));
        for (uint32_t i = tid; i < LZB_FINW; i += LZB_THREADS) S.fin[i] = 0xffffffffu;
        __syncthreads();
        for (uint32_t m0 = 0; m0 < total; m0 += LZB_THREADS) {
            const uint32_t m = m0 + tid;
            bool open = false;
            uint32_t dst = 0, len = 0, dist = 1;
            if (m < total) {
                // word that holds set bit number m: largest wi with pref[wi] <= m
                uint32_t wi = 0;
#pragma unroll
                for (int step = 64; step; step >>= 1) if (S.pref[wi + step] <= m) wi += step;
                const uint32_t bit = __fns(S.start[wi], 0, (int)(m - S.pref[wi] + 1));
                dst = c0 + wi * 32 + bit;
                const uint8_t* rec = buf + (dst - c0);
                if (FMT == LZR_DEFLATE) { len = (uint32_t)rec[0] + 3; dist = ((uint32_t)rec[1] | ((uint32_t)rec[2] << 8)) + 1; }
                else { len = (uint32_t)rec[0] | ((uint32_t)rec[1] << 8); dist = (uint32_t)rec[2] | ((uint32_t)rec[3] << 8); }
                // a record that would read before the stream or write past it cannot come from an accepted stream: skip it
                open = len != 0 && dist != 0 && dist <= dst && dst + len <= n;
            }
            if (open) lzb_range_mark<false>(S.fin, dst - c0, min(dst + len, c0 + LZC_SPAN) - c0);
            __syncthreads();
            const uint32_t s = dst - dist;
            const uint32_t src_end = s + min(len, dist);
            const uint32_t need_a = max(s, c0) - c0, need_b = src_end > c0 ? src_end - c0 : 0;   // source bits to check
            for (;;) {
                const bool ready = open && (need_b <= need_a || lzb_range_final(S.fin, need_a, need_b));
                if (ready && len <= 32) {
                    uint8_t* d8 = buf + (dst - c0);
                    if (dist >= len) {
                        // no self-overlap: independent loads (buffer, or output before the chunk), 8 at a time
                        for (uint32_t k0 = 0; k0 < len; k0 += 8) {
                            uint8_t t[8];
#pragma unroll
                            for (uint32_t k = 0; k < 8; ++k) if (k0 + k < len) { const uint32_t a = s + k0 + k; t[k] = a >= c0 ? buf[a - c0] : __ldcg(out + a); }
#pragma unroll
                            for (uint32_t k = 0; k < 8; ++k) if (k0 + k < len) d8[k0 + k] = t[k];
                        }
                    } else {
                        // overlapping copy: byte k equals source byte k mod dist
                        uint32_t j = 0;
                        for (uint32_t k = 0; k < len; ++k) { const uint32_t a = s + j; d8[k] = a >= c0 ? buf[a - c0] : __ldcg(out + a); j = j + 1 == dist ? 0 : j + 1; }
                    }
                }
                // long matches: the warp copies them one after the other (x % dist: every byte comes from the dist bytes
                // before the destination). The part of a destination beyond the buffer (LZ4 only) goes straight to the
                // output: no later match of this chunk can start there, and later chunks load it as final bytes.
                uint32_t lm = __ballot_sync(0xffffffffu, ready && len > 32);
                while (lm) {
                    const int l = __ffs(lm) - 1; lm &= lm - 1;
                    const uint32_t d = __shfl_sync(0xffffffffu, dst, l), ln = __shfl_sync(0xffffffffu, len, l);
                    const uint32_t di = __shfl_sync(0xffffffffu, dist, l), ss = d - di;
                    const bool ov = di < ln;
                    for (uint32_t x = lane; x < ln; x += 32) {
                        const uint32_t a = ss + (ov ? x % di : x);
                        const uint8_t v = a >= c0 ? buf[a - c0] : __ldcg(out + a);
                        if (d + x < c0 + LZC_SPAN) buf[d - c0 + x] = v; else out[d + x] = v;
                    }
                    __syncwarp();
                }
                if (ready) {
                    __threadfence_block();                            // the bytes are visible before the bits that announce them
                    lzb_range_mark<true>(S.fin, dst - c0, min(dst + len, c0 + LZC_SPAN) - c0);
                    open = false;
                }
                if (!__syncthreads_or(open ? 1 : 0)) break;
            }
        }
        // ---- write back [c0, cend): resolved bytes, untouched literals, and the still-parked records of the next chunk
        {
            const uint32_t nfull = span & ~15u;
            for (uint32_t i = tid * 16; i < nfull; i += LZB_THREADS * 16) *(uint4*)(gbase + i) = *(const uint4*)(S.buf + i);
            if (tid < 16 && nfull + tid < span) gbase[nfull + tid] = S.buf[nfull + tid];       // at most 15 tail bytes, never past n
        }
        __threadfence_block();
    }
    __syncthreads();
}

} // namespace gb
