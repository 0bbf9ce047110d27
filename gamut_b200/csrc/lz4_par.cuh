// lz4_par.cuh -- chunk-parallel walk of one LZ4 block (included by qoix.cu after lz4_write_walk).
//
// A QOIX file holds ONE LZ4 block (LZ4_decompress_fast, lz4.d:760-963): a chain of sequences (token, literal run,
// offset, match length) in which the position of every token depends on all sequences before it. The chain
// synchronises itself -- a walk started on an arbitrary byte soon lands on a true token position and follows the
// true chain from there -- so the block is cut into 256-byte chunks, one THREAD each:
//   lz4_sync_kernel    one CTA = 248 consecutive chunks (+ 8 warm-up chunks of its predecessor), the slice of the block
//                      in shared memory. Every thread walks its chunk from its first byte, then the CTA relaxes: chunks
//                      whose predecessor's exit differs from the entry they used walk again, until nothing changes
//                      (chunk 0 starts on the first token, so the fixed point is the serial walk). Per chunk: exit
//                      position, bytes produced, FINAL / ERR flags;
//   lz4_repair_kernel  the entry of each CTA's first own chunk against the true exit before it (walks forward where they
//                      differ); whatever still disagrees afterwards sends the image to the one-warp walk;
//   lz4_scan_kernel    output offsets per chunk; the image takes the parallel path only if the chain ends in a final
//                      sequence and the byte counts add up to the declared size;
//   lz4_pwrite_kernel  every chunk walks once more from its true entry with its output offset: short literal runs are
//                      copied by their lane, long ones by the whole warp, matches are parked as 4-byte records + a
//                      bitmap flag for lz4_resolve_kernel, exactly like the one-warp walk does.
// Anything inconsistent sends the image to lz4_parse_kernel (lz4_write_walk), which owns the reference's error
// semantics.
#pragma once

constexpr int LZP_CHUNK = 256;                   // bytes of LZ4 input per thread
constexpr int LZP_CTA = 256;                     // threads (= chunk slots) per CTA
constexpr int LZP_WARM = 8;
constexpr int LZP_OWN = LZP_CTA - LZP_WARM;
constexpr int LZP_WIN_WORDS = LZP_CTA * LZP_CHUNK / 4 + 64;      // staged slice (+ what header bytes may run past it)
constexpr uint32_t LZP_MIN_LEN = 4096;           // shorter blocks: one warp
constexpr size_t LZP_SMEM = sizeof(uint32_t) * LZP_WIN_WORDS;
enum { LZ4F_FINAL = 1, LZ4F_ERR = 2 };
struct __align__(16) Lz4Chunk { uint32_t exit, nout, flags, out_off; };

template <int WHICH>            // 1: scta_base, 2: wcta_base
__device__ __forceinline__ int lz4_find_job(const Lz4Job* __restrict__ jobs, int n, uint32_t c)
{
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((WHICH == 1 ? jobs[mid].scta_base : jobs[mid].wcta_base) <= c) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Bytes of the block: a window [w0, w0 + 4 * LZP_WIN_WORDS) in shared memory (word k of the window at
// [k / 32][(k + k / 64) % 32]: the lanes of a warp, 256 bytes apart, hit different banks), global memory elsewhere.
struct Lz4Window {
    const uint32_t* s; const uint8_t* in; uint32_t w0;
    __device__ __forceinline__ static uint32_t swz(uint32_t k) { return (k & ~31u) | ((k + (k >> 6)) & 31u); }
    __device__ __forceinline__ uint32_t operator()(uint32_t k) const
    {
        const uint32_t i = k - w0;
        if (i < (uint32_t)LZP_WIN_WORDS * 4u) return (s[swz(i >> 2)] >> ((i & 3u) * 8u)) & 0xffu;
        return in[k];
    }
};
struct Lz4GlobalBytes { const uint8_t* in; __device__ __forceinline__ uint32_t operator()(uint32_t k) const { return in[k]; } };

// stages block bytes [w0, w0 + window) (w0 may be "negative": the first CTA's warm-up slots); bytes outside the block are 0
__device__ __forceinline__ void lz4_stage(uint32_t* s_win, const Lz4Job& J, long long w0, int tid)
{
    const uint8_t* gbase = (const uint8_t*)((uintptr_t)J.in & ~(uintptr_t)3);
    const uint32_t a0 = (uint32_t)(J.in - gbase);
    for (int k = tid; k < LZP_WIN_WORDS; k += LZP_CTA) {
        const long long b = w0 + (long long)k * 4;                 // block offset of the word's first byte
        uint32_t v = 0;
        if (b >= 0 && b + 8 <= (long long)J.in_len) {
            const uint32_t* g = (const uint32_t*)(gbase + b);      // a0 + b .. : two aligned words, funnel-shifted
            v = __funnelshift_r(__ldg(g), __ldg(g + 1), a0 * 8);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) { const long long q = b + i; if (q >= 0 && q < (long long)J.in_len) v |= (uint32_t)J.in[q] << (8 * i); }
        }
        s_win[Lz4Window::swz((uint32_t)k)] = v;
    }
}

// Counting walk of the sequences that START in [p, limit): exit position, bytes produced, flags.
template <class RB>
__device__ __forceinline__ void lz4_count_chunk(const RB& rb, uint32_t in_len, uint32_t& p_io, uint32_t limit, uint32_t& nout_out, uint32_t& flags_out)
{
    uint32_t p = p_io, nout = 0, flags = 0;
    while (p < limit) {
        const uint32_t token = rb(p++);
        uint32_t L = token >> 4;
        if (L == 15) {
            uint32_t s2;
            do { if (p >= in_len) { flags |= LZ4F_ERR; break; } s2 = rb(p++); L += s2; } while (s2 == 255 && L < 0x7fffff00u);
            if (flags & LZ4F_ERR) break;
        }
        if (L > in_len - p) { flags |= LZ4F_ERR; break; }
        p += L; nout += L;
        if (p == in_len) { flags |= LZ4F_FINAL; break; }      // the final sequence has no match part
        if (in_len - p < 2) { flags |= LZ4F_ERR; break; }
        p += 2;
        uint32_t M = token & 15;
        if (M == 15) {
            uint32_t s2;
            do { if (p >= in_len) { flags |= LZ4F_ERR; break; } s2 = rb(p++); M += s2; } while (s2 == 255 && M < 0x7fffff00u);
            if (flags & LZ4F_ERR) break;
        }
        nout += M + 4;
        if (nout > 0x7fffffffu) { flags |= LZ4F_ERR; break; }
    }
    p_io = p; nout_out = nout; flags_out = flags;
}

__global__ void __launch_bounds__(LZP_CTA)
lz4_sync_kernel(const Lz4Job* __restrict__ jobs, int njobs, Lz4Chunk* __restrict__ chunks, uint32_t* __restrict__ entry_used)
{
    extern __shared__ __align__(16) uint32_t lzp_smem[];
    uint32_t* const s_win = lzp_smem;
    __shared__ uint32_t s_entry[LZP_CTA], s_exit[LZP_CTA], s_nout[LZP_CTA], s_todo_entry[LZP_CTA];
    __shared__ uint16_t s_todo[LZP_CTA];
    __shared__ uint8_t s_flags[LZP_CTA];
    __shared__ uint32_t s_wcount[LZP_CTA / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Lz4Job& J = jobs[lz4_find_job<1>(jobs, njobs, blockIdx.x)];
    const uint32_t local_cta = blockIdx.x - J.scta_base;
    const int lc_first = (int)(local_cta * LZP_OWN) - LZP_WARM;
    const long long w0 = (long long)lc_first * LZP_CHUNK;
    lz4_stage(s_win, J, w0, tid);
    const Lz4Window rb{s_win, J.in, (uint32_t)w0};
    const uint32_t in_len = J.in_len;
    const int lc = lc_first + tid;
    const bool active = lc >= 0 && (uint32_t)lc < J.nchunks;
    // a walk that stopped (final sequence, malformed data) leaves a neutral exit, the next chunk's own first byte:
    // after a true end the later chunks do not matter, after a false one the successor keeps its own guess
    auto walk = [&](int slc, uint32_t entry, uint32_t& x, uint32_t& n, uint32_t& f) {
        uint32_t p = entry;
        lz4_count_chunk(rb, in_len, p, min((uint32_t)(slc + 1) * (uint32_t)LZP_CHUNK, in_len), n, f);
        x = f ? (uint32_t)(slc + 1) * (uint32_t)LZP_CHUNK : p;
    };
    __syncthreads();
    {
        const uint32_t entry = active ? (uint32_t)lc * (uint32_t)LZP_CHUNK : 0u;
        uint32_t x = 0xffffffffu, n = 0, f = 0;
        if (active) walk(lc, entry, x, n, f);
        s_entry[tid] = entry; s_exit[tid] = x; s_nout[tid] = n; s_flags[tid] = (uint8_t)f;
    }
    const bool chained = tid > 0 && active && lc > 0;
    for (;;) {
        __syncthreads();
        bool stale = false; uint32_t prev = 0;
        if (chained) { prev = s_exit[tid - 1]; stale = prev != s_entry[tid]; }
        const uint32_t bal = __ballot_sync(0xffffffffu, stale);
        if (lane == 0) s_wcount[warp] = __popc(bal);
        __syncthreads();
        uint32_t off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < LZP_CTA / 32; ++w) { const uint32_t c = s_wcount[w]; off += w < warp ? c : 0; total += c; }
        if (total == 0) break;
        if (stale) { const uint32_t q = off + __popc(bal & ((1u << lane) - 1)); s_todo[q] = (uint16_t)tid; s_todo_entry[q] = prev; }
        __syncthreads();
        if ((uint32_t)tid < total) {
            const int s = s_todo[tid];
            const uint32_t entry = s_todo_entry[tid];
            uint32_t x, n, f;
            walk(lc_first + s, entry, x, n, f);
            s_entry[s] = entry; s_exit[s] = x; s_nout[s] = n; s_flags[s] = (uint8_t)f;
        }
    }
    if (tid >= LZP_WARM && active) *(uint4*)(chunks + J.chunk_base + (uint32_t)lc) = make_uint4(s_exit[tid], s_nout[tid], s_flags[tid], 0u);
    if (tid == LZP_WARM) entry_used[blockIdx.x] = s_entry[tid];
}

// mode 0: repair (walk forward from a wrong boundary through the CTA's own range); mode 1: images with a boundary that
// still disagrees lose the parallel path.
__global__ void __launch_bounds__(64)
lz4_repair_kernel(const Lz4Job* __restrict__ jobs, int njobs, uint32_t total_ctas, Lz4Chunk* chunks, uint32_t* entry_used, int mode, int* par_bad)
{
    const uint32_t cta = blockIdx.x * blockDim.x + threadIdx.x;
    if (cta >= total_ctas) return;
    const int j = lz4_find_job<1>(jobs, njobs, cta);
    const Lz4Job& J = jobs[j];
    const uint32_t local_cta = cta - J.scta_base;
    if (local_cta == 0) return;
    const uint32_t lc0 = local_cta * LZP_OWN;
    if (lc0 >= J.nchunks) return;
    Lz4Chunk* const ch = chunks + J.chunk_base;
    const uint32_t truth = __ldcg(&ch[lc0 - 1].exit);
    if (truth == entry_used[cta]) return;
    if (mode == 1) { par_bad[j] = 1; return; }
    entry_used[cta] = truth;
    const Lz4GlobalBytes rb{J.in};
    const uint32_t lc_end = min(lc0 + (uint32_t)LZP_OWN, J.nchunks);
    uint32_t p = truth;
    for (uint32_t lc = lc0; lc < lc_end; ++lc) {
        uint32_t n, f;
        lz4_count_chunk(rb, J.in_len, p, min((lc + 1) * (uint32_t)LZP_CHUNK, J.in_len), n, f);
        if (f) p = (lc + 1) * (uint32_t)LZP_CHUNK;
        const uint32_t old = __ldcg(&ch[lc].exit);
        __stcg(&ch[lc].exit, p); ch[lc].nout = n; ch[lc].flags = f;
        if (old == p) break;
    }
}

// one warp per image: the chain ends at the first chunk whose walk saw the final sequence (or an error); chunks after
// it are dead (out_off = 0xffffffff). The image takes the parallel path only if everything adds up.
__global__ void __launch_bounds__(128)
lz4_scan_kernel(const Lz4Job* __restrict__ jobs, int njobs, Lz4Chunk* chunks, const int* __restrict__ par_bad, int* par_ok)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= njobs) return;
    const Lz4Job J = jobs[warp];
    if (!J.nchunks) { if (lane == 0) par_ok[warp] = 0; return; }
    Lz4Chunk* C = chunks + J.chunk_base;
    uint32_t run = 0;
    bool ended = false, final_ok = false;
    for (uint32_t base = 0; base < J.nchunks; base += 32) {
        const uint32_t i = base + lane;
        const bool in = i < J.nchunks;
        const uint32_t fl = in ? C[i].flags : 0u;
        const uint32_t stopm = __ballot_sync(0xffffffffu, in && fl != 0);
        const int firststop = stopm ? __ffs(stopm) - 1 : 32;
        const bool live = in && !ended && lane <= firststop;
        const uint32_t n = live ? C[i].nout : 0u;
        uint32_t incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        if (in) C[i].out_off = live ? run + incl - n : 0xffffffffu;
        const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
        run = (run + tot < run) ? 0xffffffffu : run + tot;        // saturate: cannot equal a valid size again
        if (!ended && stopm) { ended = true; final_ok = __shfl_sync(0xffffffffu, fl, firststop) == LZ4F_FINAL; }
        if (run == 0xffffffffu) break;
    }
    if (lane == 0) par_ok[warp] = (!par_bad[warp] && ended && final_ok && run == J.orig && J.orig != 0) ? 1 : 0;
}

// The write walk of one chunk per thread, one sequence per lane and round. Same checks, same match records and the
// same literal bytes as lz4_write_walk; a violated check drops the image to the one-warp walk.
__global__ void __launch_bounds__(LZP_CTA)
lz4_pwrite_kernel(const Lz4Job* __restrict__ jobs, int njobs, const Lz4Chunk* __restrict__ chunks, int* par_ok)
{
    extern __shared__ __align__(16) uint32_t lzp_smem[];
    uint32_t* const s_win = lzp_smem;
    const int tid = threadIdx.x, lane = tid & 31;
    const int j = lz4_find_job<2>(jobs, njobs, blockIdx.x);
    if (!par_ok[j]) return;
    const Lz4Job& J = jobs[j];
    const uint32_t lc0 = (blockIdx.x - J.wcta_base) * LZP_CTA;
    const uint32_t w0 = lc0 * (uint32_t)LZP_CHUNK;
    lz4_stage(s_win, J, (long long)w0, tid);
    const Lz4Window rb{s_win, J.in, w0};
    const uint32_t in_len = J.in_len, orig = J.orig;
    const uint32_t lc = lc0 + tid;
    bool alive = lc < J.nchunks;
    uint32_t p = 0, o = 0;
    if (alive) {
        const Lz4Chunk c = chunks[J.chunk_base + lc];
        o = c.out_off; alive = o != 0xffffffffu;
        p = lc ? chunks[J.chunk_base + lc - 1].exit : 0u;
    }
    const uint32_t limit = min((lc + 1) * (uint32_t)LZP_CHUNK, in_len);
    alive = alive && p < limit;
    bool ok = true;
    __syncthreads();
    while (__any_sync(0xffffffffu, alive)) {
        uint32_t lit = 0, L = 0, dst = 0, M = 0, off = 0;
        if (alive) {
            do {        // one sequence; `break` = malformed
                ok = false;
                const uint32_t token = rb(p++);
                L = token >> 4;
                if (L == 15) {
                    uint32_t s2; bool bad = false;
                    do { if (p >= in_len) { bad = true; break; } s2 = rb(p++); L += s2; } while (s2 == 255 && L < 0x7fffff00u);
                    if (bad) break;
                }
                if (L > orig - o || L > in_len - p) break;
                const bool last = (uint64_t)o + L + 8 > orig;     // cpy > oend - COPYLENGTH
                if (last && o + L != orig) break;
                lit = p; dst = o;
                p += L; o += L;
                if (last) { alive = false; ok = true; break; }
                if (in_len - p < 2) break;
                off = rb(p) | (rb(p + 1) << 8);
                p += 2;
                if (off == 0 || off > o) break;
                M = token & 15;
                if (M == 15) {
                    uint32_t s2; bool bad = false;
                    do { if (p >= in_len) { bad = true; break; } s2 = rb(p++); M += s2; } while (s2 == 255 && M < 0x7fffff00u);
                    if (bad) break;
                }
                M += 4;
                if (M > orig - o || (uint64_t)o + M + 5 > orig) break;   // last 5 bytes are literals
                ok = true;
            } while (false);
            if (!ok) { alive = false; L = 0; M = 0; }
        }
        // literal runs: short ones by their lane, long ones by the whole warp (coalesced byte stores)
        uint32_t big = __ballot_sync(0xffffffffu, L >= 8);
        if (L && L < 8) for (uint32_t i = 0; i < L; ++i) J.out[dst + i] = (uint8_t)rb(lit + i);
        while (big) {
            const int l = __ffs(big) - 1; big &= big - 1;
            const uint32_t rl = __shfl_sync(0xffffffffu, lit, l), rd = __shfl_sync(0xffffffffu, dst, l), rn = __shfl_sync(0xffffffffu, L, l);
            for (uint32_t i = lane; i < rn; i += 32) J.out[rd + i] = (uint8_t)rb(rl + i);
        }
        // the match: 4-byte record + bitmap flag; pieces of at most 65535 bytes, each at least 4
        if (M) {
            uint32_t at = o, left = M;
            while (left) {
                uint32_t n = left > 65535 ? 65531 : left;
                if (left - n > 0 && left - n < 4) n -= 4;
                J.out[at] = (uint8_t)n; J.out[at + 1] = (uint8_t)(n >> 8);
                J.out[at + 2] = (uint8_t)off; J.out[at + 3] = (uint8_t)(off >> 8);
                atomicOr(J.bitmap + (at >> 5), 1u << (at & 31));
                at += n; left -= n;
            }
            o += M;
        }
        if (alive && p >= limit) alive = false;            // the next sequence starts in a later chunk
    }
    if (!ok) par_ok[j] = 0;
}
