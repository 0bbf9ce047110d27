// inflate_par.cuh -- block-parallel DEFLATE decode for batches of long streams (PNG IDAT at BASELINE configs[2]).
//
// The one-warp-per-stream decoder of inflate.cuh is bound by the latency of one serial symbol chain per stream:
// 1024 streams keep 7 of 64 warp slots per SM busy. This pipeline finds the parallelism inside a stream:
//
//   1. find     every bit position of every stream is tested for a dynamic-Huffman block header (BTYPE=2, HLIT and
//               HDIST <= 29, complete code-length code: bit-parallel masks + a 512-entry Kraft LUT); survivors
//               (about 1 in 2300 positions) are queued;
//   2. verify   one thread per queued position decodes the code lengths and keeps positions whose literal/length and
//               distance codes are complete (false positives: < 1 in 5e8 positions); kept positions land in a
//               per-stream slot array (one slot per 1024 input bits), which orders them for free;
//   3. compact  slots -> per-stream block list; block k's segment ends where block k+1's header starts;
//   4. count    one warp per candidate block: build its tables, cut the segment into 32 sub-chunks, every lane decodes
//               its sub-chunk from a guessed bit position; Huffman codes self-synchronise, so a lane's END is almost
//               always right even when its start was wrong, and re-decoding from the previous lane's end (fix-point
//               iteration, typically two rounds) gives every lane its true start and its output byte count;
//   5. walk     one warp per stream follows the chain start -> block end -> next block ...; a block whose start is a
//               counted candidate is accepted with its output offset (running sum); anything else (stored / fixed
//               blocks, missed or false candidates) is decoded on the spot by lane 0;
//   6. write    one warp per accepted block: lanes decode their sub-chunks again and write literals to their final
//               positions; an LZ77 match cannot be copied yet (its source may belong to a block that is being written
//               concurrently), so its (length, distance) is parked in the first 3 of the >= 3 output bytes it will
//               overwrite and its start is flagged in a bitmap (1 bit per output byte);
//   7. resolve  one warp per stream scans the bitmap and performs the copies in stream order, 32 matches at a time
//               wherever the sources lie before the first destination of the group.
//
// A stream for which any stage reports an inconsistency (corrupt data, output larger than the buffer, ...) is decoded
// again by the serial decoder, which owns the error semantics (miniz as called from stbdec.d:1267-1321); the fast
// path only ever accepts streams it decoded completely, so accepted results are identical to the serial decoder's.
#pragma once
#include "inflate.cuh"
#include "lz_resolve.cuh"

namespace gb {

constexpr uint32_t INFP_NONE = 0xffffffffu;
constexpr uint32_t INFP_SLOT_BITS = 1024;
constexpr uint32_t INFP_FINAL_WINDOW = 1u << 19;    // BFINAL=1 headers are searched only this close to the end (bits)
constexpr uint32_t INFP_MIN_CHUNK = 256;            // minimum sub-chunk (bits)
constexpr uint32_t INFP_TILE_WORDS = 1024;          // input words per CTA of the find kernel
enum { INFP_CTR_WORK = 0, INFP_CTR_A = 1, INFP_CTR_B = 2, INFP_CTR_Q = 3, INFP_CTR_Q2 = 4 };
enum { BLK_NEW = 0, BLK_EOB = 1, BLK_NOEOB = 2, BLK_ERR = 3 };

struct InfBlock {
    uint32_t bitpos, seg_end, end_bit, nbytes, out_off;
    uint8_t status, bfinal, is_true, pad;
    uint32_t pad2[2];
    uint32_t lane_start[32];
    uint32_t lane_bytes[32];
};
static_assert(sizeof(InfBlock) == 288, "InfBlock layout");

struct InfPar {
    uint32_t* slots;  uint32_t nslots;
    InfBlock* blocks; uint32_t maxblocks; uint32_t nblocks;
    uint32_t* bitmap;
    uint32_t in_bits;
    uint32_t ok;          // walk: chain followed to the final block
    uint32_t fail;        // write: inconsistency found later
    uint32_t eligible;    // host: the fast path is attempted for this stream
};

// Independent bit reader of one lane. The input is fetched as 16-byte vectors one vector ahead of use, so that the
// load latency (the L1 is mostly shared memory here; a miss goes to L2) never sits on the decode chain.
struct LaneReader {
    const uint4* v; uint32_t nvec; uint4 cur, nxt; uint64_t buf; int cnt; uint32_t widx;
    __device__ __forceinline__ void bind(const InflateJob& J) { v = (const uint4*)J.in; nvec = (J.in_len + 16) >> 4; }
    __device__ __forceinline__ uint4 ldv(uint32_t i) const { return i < nvec ? __ldg(v + i) : make_uint4(0, 0, 0, 0); }
    __device__ __forceinline__ uint32_t word()
    {
        const uint32_t k = widx & 3;
        const uint32_t w = k == 0 ? cur.x : k == 1 ? cur.y : k == 2 ? cur.z : cur.w;
        ++widx;
        if ((widx & 3) == 0) { cur = nxt; nxt = ldv((widx >> 2) + 1); }
        return w;
    }
    __device__ __forceinline__ void seek(uint32_t bitpos)
    {
        widx = bitpos >> 5;
        cur = ldv(widx >> 2); nxt = ldv((widx >> 2) + 1);
        const uint32_t a = word(), b = word();
        buf = (((uint64_t)b << 32) | a) >> (bitpos & 31);
        cnt = 64 - (int)(bitpos & 31);
    }
    __device__ __forceinline__ void refill() { if (cnt <= 32) { buf |= (uint64_t)word() << cnt; cnt += 32; } }
    __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)buf & ((1u << n) - 1); }
    __device__ __forceinline__ void drop(int n) { buf >>= n; cnt -= n; }
    __device__ __forceinline__ uint32_t get(int n) { uint32_t v = peek(n); drop(n); return v; }
    __device__ __forceinline__ uint32_t pos() const { return widx * 32 - (uint32_t)cnt; }
};

struct LaneRes { uint32_t end, nbytes, flags; };     // flags: 1 end-of-block seen, 2 invalid code

// Output side of one lane in write mode. Literals (and the 3-byte record of a match) are gathered into aligned 32-bit
// words. The bytes of a match after its record are a hole that the resolve pass overwrites, so they may receive
// anything: once a hole has crossed into a new word the lane owns every byte of that word that matters, and all
// stores but those of the first and last word of the lane's region are full-word stores.
// first word of a lane's region: bytes [k0, 4) only (the bytes below k0 belong to the lane before)
__device__ __noinline__ void infp_store_head(uint8_t* p, uint32_t w, uint32_t k0)
{
    for (uint32_t q = k0; q < 4; ++q) p[q] = (uint8_t)(w >> (8 * q));
}

struct LaneWriter {
    // One straight-line append for every kind of symbol (the lanes of a warp are at different symbols: a branch per
    // kind, or per byte stored, runs for a handful of lanes at a time): up to three bytes enter a 64-bit accumulator
    // at the byte position of `off`, a word that has been completed leaves with one predicated store, and a hole of
    // `n` bytes follows. Only the first word of the lane's region -- shared with the lane before -- goes byte by byte.
    uint8_t* out; uint32_t off; uint64_t acc; uint32_t k0;
    __device__ __forceinline__ void init(uint8_t* o, uint32_t at) { out = o; off = at; acc = 0; k0 = at & 3; }
    __device__ __forceinline__ void store_word(uint32_t base, uint32_t w)          // the word at base is complete (or abandoned to a hole)
    {
        if (k0 == 0) *(uint32_t*)(out + base) = w;
        else { infp_store_head(out + base, w, k0); k0 = 0; }
    }
    // k bytes of v (k <= 3), then a hole of n bytes
    __device__ __forceinline__ void append(uint32_t v, uint32_t k, uint32_t n)
    {
        acc |= (uint64_t)v << (8 * (off & 3));
        const uint32_t o1 = off + k;
        if ((o1 ^ off) & ~3u) { store_word(off & ~3u, (uint32_t)acc); acc >>= 32; }        // the bytes completed a word
        off = o1;
        const uint32_t e = off & 3;
        if (n >= 4 - e) {                                    // the hole leaves the current word: what I own of it goes out now
            if (e) store_word(off - e, (uint32_t)acc);
            acc = 0;
        }
        off += n;
    }
    __device__ __forceinline__ void flush()                  // end of the lane's region: only bytes below off are mine
    {
        const uint32_t e = off & 3, base = off & ~3u;
        for (uint32_t q = k0; q < e; ++q) out[base + q] = (uint8_t)((uint32_t)acc >> (8 * q));
        acc = 0; k0 = e;
    }
};

__device__ __forceinline__ void infp_cp4(uint32_t* smem_dst, const uint32_t* gsrc, uint32_t src_size)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(d), "l"(gsrc), "r"(src_size) : "memory");
}
__device__ __forceinline__ void infp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void infp_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

constexpr int INFP_RING = 16;       // input words per lane staged in shared memory
constexpr int INFP_MARK_WORDS = 8;  // self-synchronisation window: 256 bits after a lane's guessed start
constexpr int INFP_WARPS = 3;       // warps per CTA of the count / walk / write kernels

struct InfpWarp {                   // shared memory of one warp
    InflateSmem S;
    uint32_t ring[INFP_RING][32];
    uint32_t marks[INFP_MARK_WORDS][32];
};

// Decode modes: PLAIN counts; MARK also records the unit boundaries inside the window after `mark_base`;
// MERGE stops at the first unit boundary that is recorded in the window; STOPAT stops exactly at `stop_at`.
enum { LD_PLAIN = 0, LD_MARK = 1, LD_MERGE = 2, LD_STOPAT = 3, LD_WRITE = 4 };

// Decodes from bit `start` while the position is below `limit` (checked between lit/len units); stops after an
// end-of-block symbol. One Huffman symbol per loop iteration, literal/length and distance symbols through the same
// code path (the two tables are adjacent), so the 32 lanes of a warp stay converged although each decodes its own
// sub-chunk. The lane's input words are staged in a shared-memory ring (column `col`, word k at col[(k%16)*32]: bank
// = lane, conflict-free) by 4-byte cp.async; every lane tops its ring up in the same iteration (one in four), so the
// refill is a warp-uniform branch, and the data is needed no earlier than four iterations after it was requested.
// res.flags: 1 end-of-block, 2 invalid code, 4 (MERGE) merged with the recorded path at res.end.
template <int MODE>
__device__ __forceinline__ void lane_decode(const InflateSmem& S, uint32_t* col, const InflateJob& J, uint32_t start, uint32_t limit,
                                            LaneRes& res, uint32_t* mcol, uint32_t mark_base, uint32_t stop_at,
                                            uint8_t* out, uint32_t off, uint32_t* bitmap, uint32_t* fail)
{
    constexpr bool WRITE = MODE == LD_WRITE;
    const uint32_t* words = (const uint32_t*)J.in;
    const uint32_t nwords = (J.in_len + 8 + 3) >> 2;
    uint32_t pos = start;
    uint32_t fetched = (pos >> 5) & ~3u;
    auto fetch4 = [&]() {
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q) {
            const uint32_t idx = fetched + q;
            infp_cp4(col + (idx & (INFP_RING - 1)) * 32, words + (idx < nwords ? idx : 0), idx < nwords ? 4u : 0u);
        }
        fetched += 4;
    };
    fetch4(); fetch4(); fetch4();
    infp_commit();
    infp_wait<0>();
    const uint32_t* tab = S.lit_tab;                        // dist_tab follows at +1024
    uint32_t nbytes = 0, flags = 0, want_dist = 0, mlen = 0;
    LaneWriter W;
    if (WRITE) W.init(out, off);
    for (uint32_t it = 0; pos < limit || want_dist; ++it) {
        const uint32_t wi = pos >> 5;
        if ((it & 3) == 0) {
            if (fetched - wi <= 8) fetch4();
            infp_commit();
            infp_wait<1>();
        }
        if (MODE == LD_MARK || MODE == LD_MERGE) {
            const uint32_t d = pos - mark_base;
            if (!want_dist && d < INFP_MARK_WORDS * 32) {
                uint32_t* m = mcol + (d >> 5) * 32;
                if (MODE == LD_MARK) *m |= 1u << (d & 31);
                else if ((*m >> (d & 31)) & 1u) { flags = 4; break; }
            }
        }
        if (MODE == LD_STOPAT) { if (!want_dist && pos == stop_at) break; }
        const uint32_t w0 = col[(wi & (INFP_RING - 1)) * 32], w1 = col[((wi + 1) & (INFP_RING - 1)) * 32];
        const uint32_t bits = __funnelshift_r(w0, w1, pos);
        uint32_t e = tab[(want_dist ? (1u << INF_LIT_BITS) : 0u) + (bits & (want_dist ? (1u << INF_DIST_BITS) - 1 : (1u << INF_LIT_BITS) - 1))];
        if ((e & 15) == 0) {
            e = inf_slow(S, (int)want_dist, bits & 0x7fffu, want_dist ? INF_DIST_BITS : INF_LIT_BITS);
            if (e == 0) { flags = 2; break; }
        }
        const uint32_t len = e & 15, xb = (e >> 4) & 15, kind = (e >> 8) & 3;
        const uint32_t val = (e >> 16) + ((bits >> len) & ((1u << xb) - 1));
        pos += len + xb;
        if (kind == 3) { flags = 2; break; }
        if (!want_dist && kind == 2) { flags = 1; break; }   // end of block
        if (WRITE) {
            // literal: one byte. distance symbol: the match's 3-byte record, then the hole it will fill. length
            // symbol: nothing yet.
            const uint32_t at = W.off;
            const uint32_t rec = (mlen - 3) | ((val - 1) << 8);
            if (want_dist) {
                if (val > at) *fail = 1;                    // reaches before the start of the output
                atomicOr(bitmap + (at >> 5), 1u << (at & 31));
            }
            W.append(want_dist ? rec : (kind == 0 ? val : 0u), want_dist ? 3u : (kind == 0 ? 1u : 0u), want_dist ? mlen - 3 : 0u);
        }
        if (want_dist) want_dist = 0;
        else if (kind == 0) ++nbytes;
        else { nbytes += val; mlen = val; want_dist = 1; }
    }
    infp_wait<0>();
    if (WRITE) W.flush();
    res.end = pos; res.nbytes = nbytes; res.flags = flags;
}

__device__ __forceinline__ void infp_reader(InflateReader& R, const InflateJob& J, int lane)
{
    R.words = (const uint32_t*)J.in;
    R.nwords = (J.in_len + 8 + 3) >> 2;
    R.lane = lane;
}
__device__ __forceinline__ void infp_seek_bit(InflateReader& R, uint32_t bitpos)
{
    R.seek(bitpos >> 3);
    R.drop((int)(bitpos & 7));
    R.refill();
}
__device__ __forceinline__ uint32_t infp_bitpos(const InflateReader& R) { return R.widx * 32 - (uint32_t)R.bitcnt; }

// ---------------------------------------------------------------------------------------------
// 1. find
__global__ void __launch_bounds__(256)
infp_find_kernel(const InflateJob* jobs, const InfPar* par, const uint32_t* tile_start, int njobs,
                 uint2* vq, uint32_t vq_cap, uint32_t* ctr)
{
    __shared__ uint8_t kraft3[512];
    __shared__ int sj;
    const int tid = threadIdx.x;
    for (int x = tid; x < 512; x += 256) {
        int s = 0;
        for (int f = 0; f < 3; ++f) { int v = (x >> (3 * f)) & 7; if (v) s += 128 >> v; }
        kraft3[x] = (uint8_t)s;
    }
    if (tid == 0) {
        int lo = 0, hi = njobs;                 // last job with tile_start[j] <= blockIdx.x
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (tile_start[mid] <= blockIdx.x) lo = mid; else hi = mid; }
        sj = lo;
    }
    __syncthreads();
    const int j = sj;
    if (!par[j].eligible) return;
    const InflateJob J = jobs[j];
    const uint32_t tile = blockIdx.x - tile_start[j];
    const uint32_t* words = (const uint32_t*)J.in;
    const uint32_t nwords = (J.in_len + 8 + 3) >> 2;
    const uint32_t in_bits = J.in_len * 8;
    const uint32_t first_bit = J.parse_header ? 16u : 0u;
    auto ld = [&](uint32_t i) { return i < nwords ? __ldg(words + i) : 0u; };
    for (int it = 0; it < (int)(INFP_TILE_WORDS / 256); ++it) {
        const uint32_t wi = tile * INFP_TILE_WORDS + (uint32_t)it * 256 + (uint32_t)tid;
        const uint32_t p0 = wi * 32;
        if (p0 >= in_bits) continue;
        const uint32_t w0 = ld(wi), w1 = ld(wi + 1), w2 = ld(wi + 2), w3 = ld(wi + 3);
        auto S = [&](int k) { return __funnelshift_r(w0, w1, k); };
        uint32_t cand = ~S(1) & S(2) & ~(S(4) & S(5) & S(6) & S(7)) & ~(S(9) & S(10) & S(11) & S(12));
        if (p0 + 32 + INFP_FINAL_WINDOW < in_bits) cand &= ~w0;          // BFINAL must be 0 far from the end
        if (p0 < first_bit) cand &= ~((1u << (first_bit - p0)) - 1);       // first_bit is 0 or 16
        if (p0 + 32 + 64 > in_bits) {                                     // a block needs a header and an EOB
            for (int b = 0; b < 32; ++b) if (p0 + b + 64 > in_bits) cand &= ~(1u << b);
        }
        while (cand) {
            const int b = __ffs(cand) - 1;
            cand &= cand - 1;
            const uint32_t x0 = __funnelshift_r(w0, w1, b), x1 = __funnelshift_r(w1, w2, b), x2 = __funnelshift_r(w2, w3, b);
            const uint32_t nb = (((x0 >> 13) & 15) + 4) * 3;
            uint32_t vlo = __funnelshift_r(x0, x1, 17), vhi = __funnelshift_r(x1, x2, 17);
            if (nb < 32) { vlo &= (1u << nb) - 1; vhi = 0; } else vhi &= (1u << (nb - 32)) - 1;
            const uint32_t g3 = ((vlo >> 27) | (vhi << 5)) & 511;
            const int k = kraft3[vlo & 511] + kraft3[(vlo >> 9) & 511] + kraft3[(vlo >> 18) & 511] + kraft3[g3] +
                          kraft3[(vhi >> 4) & 511] + kraft3[(vhi >> 13) & 511] + kraft3[(vhi >> 22) & 511];
            if (k == 128) {
                const uint32_t q = atomicAdd(ctr + INFP_CTR_Q, 1u);
                if (q < vq_cap) vq[q] = make_uint2((uint32_t)j, p0 + (uint32_t)b);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 2. verify: the code lengths of the header must give complete literal/length and distance codes.
// USE_LUT: the code-length code is decoded through a 128-entry per-thread table in shared memory (worth its set-up
// cost for the headers that survive the first, short pass) instead of the canonical compare chain in registers.
template <bool USE_LUT>
__global__ void __launch_bounds__(128)
infp_verify_kernel(const InflateJob* jobs, const InfPar* par, const uint2* vq, uint32_t vq_cap, const uint32_t* qcount,
                   int max_syms, uint2* vq_next, uint32_t vq_next_cap, uint32_t* qcount_next)
{
    __shared__ uint8_t lut[USE_LUT ? 128 : 1][128];          // [7 code bits][thread]: symbol << 3 | length
    uint32_t n = *qcount;
    if (n > vq_cap) n = vq_cap;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const uint2 c = vq[q];
        const InflateJob& J = jobs[c.x];
        LaneReader R;
        R.bind(J);
        R.seek(c.y + 3);
        const int nlit = (int)R.get(5) + 257, ndist = (int)R.get(5) + 1, ncl = (int)R.get(4) + 4;
        uint32_t cl_lens_lo = 0, cl_lens_hi = 0;
        for (int i = 0; i < ncl; ++i) {
            R.refill();
            const uint32_t v = R.get(3);
            const int sidx = inf_cl_order[i];
            if (sidx < 10) cl_lens_lo |= v << (3 * sidx); else cl_lens_hi |= v << (3 * (sidx - 10));
        }
        int cl_count[8];
#pragma unroll
        for (int l = 0; l < 8; ++l) cl_count[l] = 0;
#pragma unroll
        for (int s = 0; s < 19; ++s) {
            const int l = (s < 10) ? (cl_lens_lo >> (3 * s)) & 7 : (cl_lens_hi >> (3 * (s - 10))) & 7;
#pragma unroll
            for (int k = 1; k < 8; ++k) cl_count[k] += (l == k);
        }
        int cl_first[8], cl_fsym[8];
        {
            int code = 0, sym = 0;
            cl_first[0] = 0; cl_fsym[0] = 0;
#pragma unroll
            for (int l = 1; l < 8; ++l) {
                code = (code + (l > 1 ? cl_count[l - 1] : 0)) << 1;
                cl_first[l] = code; cl_fsym[l] = sym; sym += cl_count[l];
            }
        }
        uint64_t cl_sorted_lo = 0, cl_sorted_hi = 0;
        if (USE_LUT) {
            // canonical codes in symbol order; the code is complete (checked by the find kernel), so every one of the
            // 128 patterns is covered and the table needs no clearing
            int nxt[8];
#pragma unroll
            for (int l = 0; l < 8; ++l) nxt[l] = cl_first[l];
#pragma unroll
            for (int sy = 0; sy < 19; ++sy) {
                const int l = (sy < 10) ? (cl_lens_lo >> (3 * sy)) & 7 : (cl_lens_hi >> (3 * (sy - 10))) & 7;
                if (l) {
                    int code = 0;
#pragma unroll
                    for (int k = 1; k < 8; ++k) if (l == k) { code = nxt[k]; nxt[k] = code + 1; }
                    const uint32_t rev = __brev((uint32_t)code) >> (32 - l);
                    for (uint32_t k = rev; k < 128; k += (1u << l)) lut[k][threadIdx.x] = (uint8_t)((sy << 3) | l);
                }
            }
        } else {
            int k = 0;
#pragma unroll
            for (int l = 1; l < 8; ++l) {
#pragma unroll
                for (int s = 0; s < 19; ++s) {
                    const int sl = (s < 10) ? (cl_lens_lo >> (3 * s)) & 7 : (cl_lens_hi >> (3 * (s - 10))) & 7;
                    if (sl == l) {
                        if (k < 12) cl_sorted_lo |= (uint64_t)s << (5 * k); else cl_sorted_hi |= (uint64_t)s << (5 * (k - 12));
                        ++k;
                    }
                }
            }
        }
        const int total = nlit + ndist;
        int i = 0, prev = 0, nd = 0;
        uint32_t kl = 0, kd = 0;
        bool good = true, has256 = false;
        int nsym = 0;
        while (i < total) {
            if (++nsym > max_syms) break;
            R.refill();
            int sym = -1, len = 0;
            if (USE_LUT) {
                const uint32_t e = lut[R.peek(7)][threadIdx.x];
                len = (int)(e & 7); sym = len ? (int)(e >> 3) : -1;
            } else {
                const uint32_t rev = __brev(R.peek(7)) >> 25;
#pragma unroll
                for (int l = 1; l < 8; ++l) {
                    if (sym < 0) {
                        const int cc = (int)(rev >> (7 - l));
                        const int d = cc - cl_first[l];
                        if (d >= 0 && d < cl_count[l]) {
                            const int k = cl_fsym[l] + d;
                            sym = (k < 12) ? (int)((cl_sorted_lo >> (5 * k)) & 31) : (int)((cl_sorted_hi >> (5 * (k - 12))) & 31);
                            len = l;
                        }
                    }
                }
            }
            if (sym < 0) { good = false; break; }
            R.drop(len);
            int rep = 1, val = sym;
            if (sym == 16) { if (i == 0) { good = false; break; } rep = 3 + (int)R.get(2); val = prev; }
            else if (sym == 17) { rep = 3 + (int)R.get(3); val = 0; }
            else if (sym == 18) { rep = 11 + (int)R.get(7); val = 0; }
            if (i + rep > total) { good = false; break; }
            if (val) {
                const uint32_t wgt = 32768u >> val;
                const int a = i < nlit ? min(rep, nlit - i) : 0, b = rep - a;
                kl += (uint32_t)a * wgt; kd += (uint32_t)b * wgt; nd += b;
                if (kl > 32768u || kd > 32768u) { good = false; break; }     // over-subscribed: most random headers end here
                if (i <= 256 && 256 < i + rep) has256 = true;
            }
            i += rep; prev = val;
        }
        if (good && i < total) {                // ran out of this pass's symbol budget: the next pass finishes it
            const uint32_t q2 = atomicAdd(qcount_next, 1u);
            if (q2 < vq_next_cap) vq_next[q2] = c;
        } else if (good && has256 && kl == 32768u && (kd == 32768u || nd <= 1))
            atomicMin(par[c.x].slots + c.y / INFP_SLOT_BITS, c.y);
    }
}

// ---------------------------------------------------------------------------------------------
// 3. compact: slot array -> block list (+ global work list); slots then map a position's slot to its block index.
__global__ void __launch_bounds__(128)
infp_compact_kernel(InfPar* par, int njobs, uint2* work, uint32_t* ctr)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 4 + warp;
    if (j >= njobs) return;
    InfPar& P = par[j];
    if (!P.eligible) return;
    uint32_t nb = 0;
    for (uint32_t base = 0; base < P.nslots; base += 32) {
        const uint32_t s = base + lane;
        const uint32_t v = s < P.nslots ? P.slots[s] : INFP_NONE;
        const bool has = v != INFP_NONE;
        const uint32_t m = __ballot_sync(0xffffffffu, has);
        const uint32_t idx = nb + __popc(m & ((1u << lane) - 1));
        if (has) {
            if (idx < P.maxblocks) { P.blocks[idx].bitpos = v; P.slots[s] = idx; }
            else P.slots[s] = INFP_NONE;
        }
        nb += __popc(m);
    }
    if (nb > P.maxblocks) nb = P.maxblocks;
    __syncwarp();
    for (uint32_t i = lane; i < nb; i += 32) {
        InfBlock& B = P.blocks[i];
        B.seg_end = i + 1 < nb ? P.blocks[i + 1].bitpos : P.in_bits;
        B.status = BLK_NEW; B.is_true = 0; B.bfinal = 0; B.nbytes = 0; B.end_bit = 0; B.out_off = 0;
    }
    uint32_t base = 0;
    if (lane == 0) { P.nblocks = nb; base = atomicAdd(ctr + INFP_CTR_WORK, nb); }
    base = __shfl_sync(0xffffffffu, base, 0);
    for (uint32_t i = lane; i < nb; i += 32) work[base + i] = make_uint2((uint32_t)j, i);
}

// sub-chunk geometry of a block body [body, e)
__device__ __forceinline__ void infp_chunks(uint32_t body, uint32_t e, int lane, uint32_t& g, uint32_t& limit)
{
    const uint32_t span = e > body ? e - body : 0;
    uint32_t c = (span + 31) / 32;
    if (c < INFP_MIN_CHUNK) c = INFP_MIN_CHUNK;
    const uint64_t gg = (uint64_t)body + (uint64_t)lane * c;
    g = gg < e ? (uint32_t)gg : INFP_NONE;
    const uint64_t ll = gg + c;
    limit = ll < e ? (uint32_t)ll : e;
}

// ---------------------------------------------------------------------------------------------
// 4. count. Block B of stream J with its segment end B.seg_end: builds the tables, finds every lane's true start and
// byte count. Round 0 decodes every sub-chunk from its guessed start and records the unit boundaries of the first 256
// bits; later rounds decode from the previous lane's end only until they land on a recorded boundary (Huffman codes
// self-synchronise within a few symbols), and take the rest of the sub-chunk from round 0.
__device__ inline void infp_count_block(const InflateJob& J, InfBlock& B, InfpWarp& M, int lane)
{
    InflateSmem& S = M.S;
    InflateReader R;
    infp_reader(R, J, lane);
    infp_seek_bit(R, B.bitpos);
    const uint32_t bfinal = R.get(1), btype = R.get(2);
    __syncwarp();
    if (btype != 2 || !inf_setup_tables(R, S, lane, btype)) {
        if (lane == 0) B.status = BLK_ERR;
        return;
    }
    const uint32_t body = infp_bitpos(R), e = B.seg_end;
    uint32_t g, limit;
    infp_chunks(body, e, lane, g, limit);
    uint32_t* col = &M.ring[0][lane];
    uint32_t* mcol = &M.marks[0][lane];
#pragma unroll
    for (int k = 0; k < INFP_MARK_WORDS; ++k) mcol[k * 32] = 0;
    const uint32_t g0 = lane == 0 ? body : g;              // round-0 start
    LaneRes r0; r0.end = g0; r0.nbytes = 0; r0.flags = 0;
    if (g0 != INFP_NONE) lane_decode<LD_MARK>(S, col, J, g0, limit, r0, mcol, g0, 0, nullptr, 0, nullptr, nullptr);
    uint32_t start = g0;
    LaneRes res = r0;
    for (int round = 0; round < 40; ++round) {
        const uint32_t pend = __shfl_up_sync(0xffffffffu, res.end, 1);
        const uint32_t pfl = __shfl_up_sync(0xffffffffu, res.flags, 1);
        const uint32_t pst = __shfl_up_sync(0xffffffffu, start, 1);
        uint32_t ns = (pst != INFP_NONE && pfl == 0 && pend < e) ? pend : INFP_NONE;
        if (lane == 0) ns = body;
        const bool changed = ns != start;
        if (!__any_sync(0xffffffffu, changed)) break;
        if (changed) {
            start = ns;
            res.end = start; res.nbytes = 0; res.flags = 0;
            if (start == g0) res = r0;
            else if (start != INFP_NONE) {
                LaneRes a;
                lane_decode<LD_MERGE>(S, col, J, start, limit, a, mcol, g0, 0, nullptr, 0, nullptr, nullptr);
                if (a.flags == 4) {
                    // merged with the round-0 path at a.end: bytes = mine up to there + round 0's from there on
                    LaneRes c;
                    lane_decode<LD_STOPAT>(S, col, J, g0, limit, c, nullptr, 0, a.end, nullptr, 0, nullptr, nullptr);
                    res.end = r0.end; res.flags = r0.flags; res.nbytes = a.nbytes + (r0.nbytes - c.nbytes);
                } else res = a;
            }
        }
    }
    const bool alive = start != INFP_NONE;
    const uint32_t stopm = __ballot_sync(0xffffffffu, alive && res.flags != 0);
    const uint32_t alivem = __ballot_sync(0xffffffffu, alive);
    int status = BLK_NOEOB;
    int last = 31 - __clz(alivem);            // alive lanes form a prefix (lane 0 is always alive)
    if (stopm) {
        last = __ffs(stopm) - 1;
        status = (__shfl_sync(0xffffffffu, res.flags, last) == 1) ? BLK_EOB : BLK_ERR;
    }
    uint32_t nb = lane <= last ? res.nbytes : 0;
    uint32_t tot = nb;
    bool ovf = false;
#pragma unroll
    for (int d = 16; d; d >>= 1) { const uint32_t o = __shfl_xor_sync(0xffffffffu, tot, d); ovf |= (tot + o) < tot; tot += o; }
    if (__any_sync(0xffffffffu, ovf) && status == BLK_EOB) status = BLK_ERR;
    B.lane_start[lane] = lane <= last ? start : INFP_NONE;
    B.lane_bytes[lane] = nb;
    const uint32_t endb = __shfl_sync(0xffffffffu, res.end, last);
    if (lane == 0) { B.end_bit = endb; B.nbytes = tot; B.bfinal = (uint8_t)bfinal; B.status = (uint8_t)status; }
    __syncwarp();
}

__global__ void __launch_bounds__(INFP_WARPS * 32)
infp_count_kernel(const InflateJob* jobs, InfPar* par, const uint2* work, uint32_t* ctr)
{
    __shared__ InfpWarp smem[INFP_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nwork = ctr[INFP_CTR_WORK];
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(ctr + INFP_CTR_A, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= nwork) break;
        const uint2 wk = work[i];
        infp_count_block(jobs[wk.x], par[wk.x].blocks[wk.y], smem[warp], lane);
    }
}

// ---------------------------------------------------------------------------------------------
// 5. walk
__global__ void __launch_bounds__(INFP_WARPS * 32)
infp_walk_kernel(InflateJob* jobs, InfPar* par, int njobs)
{
    __shared__ InfpWarp smem[INFP_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * INFP_WARPS + warp;
    if (j >= njobs) return;
    InfPar& P = par[j];
    if (!P.eligible) return;
    InflateSmem& S = smem[warp].S;
    uint32_t* col = &smem[warp].ring[0][lane];
    const InflateJob J = jobs[j];
    const uint32_t in_bits = P.in_bits, cap = J.out_cap;
    uint32_t pos = 0, out_off = 0;
    bool ok = false;
    if (J.parse_header) {
        if (J.in_len < 2) return;
        const uint32_t cmf = J.in[0], flg = J.in[1];
        if (((cmf * 256 + flg) % 31 != 0) || (flg & 32) || ((cmf & 15) != 8)) return;
        pos = 16;
    }
    InflateReader R;
    infp_reader(R, J, lane);
    for (;;) {
        if (pos + 3 > in_bits) break;
        const uint32_t slot = pos / INFP_SLOT_BITS;
        const uint32_t bi = slot < P.nslots ? P.slots[slot] : INFP_NONE;
        if (bi < P.nblocks) {
            InfBlock& B = P.blocks[bi];
            if (B.bitpos == pos) {
                // a false candidate inside this block cut its segment short: count it again with a later segment end
                for (uint32_t k = bi + 2; B.status == BLK_NOEOB && k <= bi + 4 && k <= P.nblocks; ++k) {
                    __syncwarp();
                    if (lane == 0) B.seg_end = k < P.nblocks ? P.blocks[k].bitpos : in_bits;
                    __syncwarp();
                    infp_count_block(J, B, smem[warp], lane);
                }
                if (B.status == BLK_EOB) {
                    const uint32_t nb = B.nbytes;
                    if (nb > cap - out_off) break;
                    const uint32_t e = B.end_bit;
                    const uint32_t fin = B.bfinal;
                    __syncwarp();
                    if (lane == 0) { B.out_off = out_off; B.is_true = 1; }
                    out_off += nb; pos = e;
                    if (fin) { ok = true; break; }
                    continue;
                }
            }
        }
        // not a counted candidate: decode this block here
        infp_seek_bit(R, pos);
        const uint32_t bfinal = R.get(1), btype = R.get(2);
        if (btype == 0) {
            R.drop(R.bitcnt & 7);
            R.refill();
            const uint32_t len = R.get(16);
            R.refill();
            const uint32_t nlen = R.get(16);
            if ((len ^ 0xffffu) != nlen) break;
            const uint32_t bp = R.bytepos_ceil();
            if (bp + len > J.in_len) break;
            if (len > cap - out_off) break;
            for (uint32_t i = lane; i < len; i += 32) J.out[out_off + i] = J.in[bp + i];
            out_off += len;
            pos = (bp + len) * 8;
        } else if (btype == 3) {
            break;
        } else {
            __syncwarp();
            if (!inf_setup_tables(R, S, lane, btype)) break;
            const uint32_t body = infp_bitpos(R);
            LaneRes res; res.end = body; res.nbytes = 0; res.flags = 0;
            const uint32_t lim = in_bits + 64 > in_bits ? in_bits + 64 : 0xffffffffu;
            if (lane == 0) lane_decode<LD_PLAIN>(S, col, J, body, lim, res, nullptr, 0, 0, nullptr, 0, nullptr, nullptr);
            res.end = __shfl_sync(0xffffffffu, res.end, 0);
            res.nbytes = __shfl_sync(0xffffffffu, res.nbytes, 0);
            res.flags = __shfl_sync(0xffffffffu, res.flags, 0);
            if (res.flags != 1) break;
            if (res.nbytes > cap - out_off) break;
            if (lane == 0) {
                LaneRes r2;
                lane_decode<LD_WRITE>(S, col, J, body, lim, r2, nullptr, 0, 0, J.out, out_off, P.bitmap, &P.fail);
            }
            __syncwarp();
            out_off += res.nbytes;
            pos = res.end;
        }
        if (bfinal) { ok = true; break; }
    }
    if (ok && pos <= in_bits && lane == 0) {
        P.ok = 1;
        jobs[j].out_len = out_off;
        jobs[j].status = INF_OK;
    }
}

// ---------------------------------------------------------------------------------------------
// 6. write
__global__ void __launch_bounds__(INFP_WARPS * 32)
infp_write_kernel(const InflateJob* jobs, InfPar* par, const uint2* work, uint32_t* ctr)
{
    __shared__ InfpWarp smem[INFP_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    InflateSmem& S = smem[warp].S;
    const uint32_t nwork = ctr[INFP_CTR_WORK];
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(ctr + INFP_CTR_B, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= nwork) break;
        const uint2 wk = work[i];
        InfPar& P = par[wk.x];
        const InfBlock& B = P.blocks[wk.y];
        if (!P.ok || !B.is_true) continue;
        const InflateJob& J = jobs[wk.x];
        InflateReader R;
        infp_reader(R, J, lane);
        infp_seek_bit(R, B.bitpos);
        R.get(1);
        const uint32_t btype = R.get(2);
        __syncwarp();
        if (!inf_setup_tables(R, S, lane, btype)) { if (lane == 0) P.fail = 1; continue; }
        const uint32_t body = infp_bitpos(R);
        uint32_t g, limit;
        infp_chunks(body, B.seg_end, lane, g, limit);
        const uint32_t start = B.lane_start[lane], nb = B.lane_bytes[lane];
        uint32_t incl = nb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
        const uint32_t off = B.out_off + incl - nb;
        if (start != INFP_NONE) {
            LaneRes res;
            lane_decode<LD_WRITE>(S, &smem[warp].ring[0][lane], J, start, limit, res, nullptr, 0, 0, J.out, off, P.bitmap, &P.fail);
            if (res.nbytes != nb) P.fail = 1;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// 7. resolve (lz_resolve.cuh)
__global__ void __launch_bounds__(LZ4C_NT)
infp_resolve_kernel(const InflateJob* jobs, const InfPar* par, int njobs)
{
    __shared__ Lz4cShared S;
    const int j = blockIdx.x;
    if (j >= njobs) return;
    const InfPar& P = par[j];
    if (!P.eligible || !P.ok || P.fail) return;
    lz_resolve_stream_jump<LZR_DEFLATE>(jobs[j].out, jobs[j].out_len, P.bitmap, S);
}

} // namespace gb
