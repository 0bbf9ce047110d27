// inflate.cuh -- device-side DEFLATE (RFC 1951) / zlib (RFC 1950) stream decoder, one warp per stream.
//
// Replaces miniz `mz_uncompress3` as called from stbi_zlib_decode_malloc_guesssize_headerflag
// (source/gamut/codecs/stbdec.d:1267-1321): zlib header checked when parse_header, Adler-32 neither
// read nor verified (trusted_input), trailing bytes tolerated.
//
// Design (B200): the byte stream is serial, so one warp owns one stream and all 32 lanes execute the
// symbol decode redundantly (warp-uniform control flow, no divergence); the lanes are used for
//   - input staging: every lane holds one 32-bit word of the current and of the next 128-byte input
//     chunk (one coalesced load per 128 B, prefetched one chunk ahead); refills are warp shuffles;
//   - Huffman tables: 10-bit (lit/len) and 9-bit (distance) lookup tables in shared memory, filled
//     cooperatively; longer codes take a canonical slow path;
//   - output: literals are batched 32 at a time (lane k keeps the k-th literal) and written with one
//     coalesced store; LZ77 matches are copied 32 bytes per step by the whole warp.
#pragma once
#include <stdint.h>

namespace gb {

enum InflateStatus { INF_OK = 0, INF_OUTPUT_FULL = 1, INF_DATA_ERROR = 2 };

struct InflateJob {
    const uint8_t* in;      // 4-byte aligned; at least in_len + 8 readable bytes
    uint32_t in_len;
    uint8_t* out;
    uint32_t out_cap;
    int parse_header;       // 1: zlib wrapper, 0: raw deflate (CgBI)
    // results
    uint32_t out_len;
    int status;
};

static __constant__ uint8_t inf_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

constexpr int INF_LIT_BITS = 10;
constexpr int INF_DIST_BITS = 9;
constexpr int INF_WARPS_PER_CTA = 4;

struct InflateSmem {
    uint32_t lit_tab[1 << INF_LIT_BITS];
    uint32_t dist_tab[1 << INF_DIST_BITS];
    uint32_t sorted_entry[320];   // decode entry | code length, in canonical (length, symbol) order
    uint8_t  lens[320];           // code lengths: 0..287 lit/len, 288..319 distance
    uint16_t rank[320];           // rank of a symbol among the symbols of its length (build scratch)
    uint16_t count[2][16];
    uint16_t first_code[2][16];
    uint16_t first_sym[2][16];
    uint32_t lj_end[2][16];       // left-justified (15-bit) end of the codes of each length
    int      err;
};

// entry layout: [3:0] code length (0 = take the slow path), [7:4] extra-bit count,
// [9:8] kind (0 literal, 1 length/distance base, 2 end-of-block, 3 invalid), [31:16] value/base
__device__ __forceinline__ uint32_t inf_entry(int kind, int extra, int value) { return (uint32_t)(extra << 4) | (uint32_t)(kind << 8) | ((uint32_t)value << 16); }

// decode entry (without code length) of symbol s; s >= 288: distance symbol s - 288 (RFC 1951 3.2.5)
__device__ __forceinline__ uint32_t inf_sym_entry(int s)
{
    if (s < 256) return inf_entry(0, 0, s);
    if (s == 256) return inf_entry(2, 0, 0);
    if (s < 265) return inf_entry(1, 0, s - 257 + 3);
    if (s < 285) { const int x = (s - 261) >> 2; return inf_entry(1, x, 3 + ((4 + ((s - 265) & 3)) << x)); }
    if (s == 285) return inf_entry(1, 0, 258);
    if (s < 288) return inf_entry(3, 0, 0);
    const int d = s - 288;
    if (d < 4) return inf_entry(1, 0, d + 1);
    if (d < 30) { const int x = (d >> 1) - 1; return inf_entry(1, x, 1 + ((2 + (d & 1)) << x)); }
    return inf_entry(3, 0, 0);
}

struct InflateReader {
    const uint32_t* words;   // input as aligned words
    uint32_t nwords;         // number of readable words (padded)
    uint64_t bitbuf;
    int bitcnt;
    uint32_t widx;           // next word to feed into bitbuf
    uint32_t cur, nxt;       // this lane's word of chunk (widx>>5) and of the next chunk
    int lane;

    __device__ __forceinline__ uint32_t ldw(uint32_t i) const { return i < nwords ? __ldg(words + i) : 0u; }
    __device__ __forceinline__ void seek(uint32_t bytepos)
    {
        widx = bytepos >> 2;
        uint32_t base = widx & ~31u;
        cur = ldw(base + lane);
        nxt = ldw(base + 32 + lane);
        bitbuf = 0; bitcnt = 0;
        refill();
        int drop = (bytepos & 3) * 8;
        bitbuf >>= drop; bitcnt -= drop;
        refill();
    }
    __device__ __forceinline__ void refill()
    {
        if (bitcnt <= 32) {
            uint32_t w = __shfl_sync(0xffffffffu, cur, widx & 31);
            bitbuf |= (uint64_t)w << bitcnt;
            bitcnt += 32;
            ++widx;
            if ((widx & 31) == 0) { cur = nxt; nxt = ldw(widx + 32 + lane); }
        }
    }
    __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)bitbuf & ((1u << n) - 1); }
    __device__ __forceinline__ void drop(int n) { bitbuf >>= n; bitcnt -= n; }
    __device__ __forceinline__ uint32_t get(int n) { uint32_t v = peek(n); drop(n); return v; }
    // bytes consumed so far (bits still buffered are not consumed)
    __device__ __forceinline__ uint32_t bytepos_ceil() const { return widx * 4 - (uint32_t)(bitcnt >> 3); }
};

// Builds decode tables for one code (which = 0 lit/len with n<=288 symbols at lens[0..], 1 distance
// at lens[288..]). Warp-cooperative: ranks by __match_any_sync, 16-step canonical prefix on lane 0, table fill
// by all lanes. Returns false on an over-subscribed code.
__device__ inline bool inf_build(InflateSmem& S, int which, int n, int lane)
{
    const int off = which ? 288 : 0;
    const int FAST = which ? INF_DIST_BITS : INF_LIT_BITS;
    uint32_t* tab = which ? S.dist_tab : S.lit_tab;
    for (int i = lane; i < (1 << FAST); i += 32) tab[i] = 0;
    if (lane < 16) S.count[which][lane] = 0;
    __syncwarp();
    for (int base = 0; base < n; base += 32) {
        const int s = base + lane;
        const int l = s < n ? S.lens[off + s] : 0;
        const uint32_t m = __match_any_sync(0xffffffffu, l);
        const int before = S.count[which][l];
        __syncwarp();
        if (s < n) S.rank[off + s] = (uint16_t)(before + __popc(m & ((1u << lane) - 1)));
        if ((int)(__ffs(m) - 1) == lane) S.count[which][l] = (uint16_t)(before + __popc(m));
        __syncwarp();
    }
    if (lane == 0) {
        S.count[which][0] = 0;
        int code = 0, sym = 0, left = 1;
        bool over = false;
        S.lj_end[which][0] = 0;
        for (int l = 1; l < 16; ++l) {
            left <<= 1;
            left -= S.count[which][l];
            if (left < 0) over = true;
            code = (code + S.count[which][l - 1]) << 1;
            S.first_code[which][l] = (uint16_t)code;
            S.first_sym[which][l] = (uint16_t)sym;
            sym += S.count[which][l];
            S.lj_end[which][l] = (uint32_t)(code + S.count[which][l]) << (15 - l);
        }
        S.err = over ? 1 : 0;
    }
    __syncwarp();
    if (S.err) return false;
    for (int s = lane; s < n; s += 32) {
        const int l = S.lens[off + s];
        if (l > 0) {
            const uint32_t r = S.rank[off + s];
            const uint32_t e = inf_sym_entry(off + s) | (uint32_t)l;
            S.sorted_entry[off + S.first_sym[which][l] + r] = e;
            if (l <= FAST) {
                const uint32_t rev = __brev((uint32_t)S.first_code[which][l] + r) >> (32 - l);
                for (uint32_t k = rev; k < (1u << FAST); k += (1u << l)) tab[k] = e;
            }
        }
    }
    __syncwarp();
    return true;
}

// Slow path: canonical decode of a code longer than FAST bits. Returns entry|len or 0 (invalid).
// Codes of length l, left-justified to 15 bits, fill [lj_end[l-1], lj_end[l]).
__device__ __forceinline__ uint32_t inf_slow(const InflateSmem& S, int which, uint32_t bits15, int FAST)
{
    const uint32_t rev = __brev(bits15) >> 17;   // 15 bits, first-read bit is the MSB
    int l = FAST + 1;
#pragma unroll
    for (int k = 10; k < 15; ++k) if (k > FAST) l += rev >= S.lj_end[which][k];
    if (rev >= S.lj_end[which][l]) return 0;
    const uint32_t d = (rev - S.lj_end[which][l - 1]) >> (15 - l);
    return S.sorted_entry[(which ? 288 : 0) + S.first_sym[which][l] + d];
}

// Reads the code tables of a Huffman block (btype 1: fixed, 2: dynamic; the 3 header bits are already
// consumed) and builds the decode tables. Warp-uniform. Returns false on a malformed header.
__device__ inline bool inf_setup_tables(InflateReader& R, InflateSmem& S, int lane, uint32_t btype)
{
    int nlit, ndist;
    if (btype == 1) {
        for (int s = lane; s < 288; s += 32) S.lens[s] = (s < 144) ? 8 : (s < 256) ? 9 : (s < 280) ? 7 : 8;
        S.lens[288 + lane] = 5;
        nlit = 288; ndist = 32;
        __syncwarp();
    } else {
        R.refill();
        nlit = (int)R.get(5) + 257;
        ndist = (int)R.get(5) + 1;
        int ncl = (int)R.get(4) + 4;
        // code-length code: 19 symbols, 3 bits each, in the RFC's permuted order
        uint32_t cl_lens_lo = 0, cl_lens_hi = 0;           // 19 x 3 bits packed (symbol-indexed)
        for (int i = 0; i < ncl; ++i) {
            R.refill();
            uint32_t v = R.get(3);
            int sidx = inf_cl_order[i];
            if (sidx < 10) cl_lens_lo |= v << (3 * sidx); else cl_lens_hi |= v << (3 * (sidx - 10));
        }
        // tiny canonical decoder for the code-length code (max 7 bits), in registers
        int cl_count[8];
#pragma unroll
        for (int l = 0; l < 8; ++l) cl_count[l] = 0;
#pragma unroll
        for (int s = 0; s < 19; ++s) {
            int l = (s < 10) ? (cl_lens_lo >> (3 * s)) & 7 : (cl_lens_hi >> (3 * (s - 10))) & 7;
#pragma unroll
            for (int q = 1; q < 8; ++q) cl_count[q] += (l == q);
        }
        int cl_first[8], cl_fsym[8];
        {
            int code = 0, sym = 0, left = 1; bool over = false;
            cl_first[0] = 0; cl_fsym[0] = 0;
#pragma unroll
            for (int l = 1; l < 8; ++l) {
                left = (left << 1) - cl_count[l];
                if (left < 0) over = true;
                code = (code + (l > 1 ? cl_count[l - 1] : 0)) << 1;
                cl_first[l] = code; cl_fsym[l] = sym; sym += cl_count[l];
            }
            if (over) return false;
        }
        // sorted symbol list of the code-length code, packed 5 bits each into two 64-bit words
        uint64_t cl_sorted_lo = 0, cl_sorted_hi = 0;
        {
            int k = 0;
#pragma unroll
            for (int l = 1; l < 8; ++l) {
#pragma unroll
                for (int s = 0; s < 19; ++s) {
                    int sl = (s < 10) ? (cl_lens_lo >> (3 * s)) & 7 : (cl_lens_hi >> (3 * (s - 10))) & 7;
                    if (sl == l) {
                        if (k < 12) cl_sorted_lo |= (uint64_t)s << (5 * k); else cl_sorted_hi |= (uint64_t)s << (5 * (k - 12));
                        ++k;
                    }
                }
            }
        }
        int total = nlit + ndist;
        int i = 0, prev = 0;
        while (i < total) {
            R.refill();
            uint32_t rev = __brev(R.peek(7)) >> 25;
            int sym = -1, len = 0;
#pragma unroll
            for (int l = 1; l < 8; ++l) {
                if (sym < 0) {
                    int c = (int)(rev >> (7 - l));
                    int d = c - cl_first[l];
                    if (d >= 0 && d < cl_count[l]) {
                        int k = cl_fsym[l] + d;
                        sym = (k < 12) ? (int)((cl_sorted_lo >> (5 * k)) & 31) : (int)((cl_sorted_hi >> (5 * (k - 12))) & 31);
                        len = l;
                    }
                }
            }
            if (sym < 0) return false;
            R.drop(len);
            int rep = 1, val = sym;
            if (sym == 16) { if (i == 0) return false; rep = 3 + (int)R.get(2); val = prev; }
            else if (sym == 17) { rep = 3 + (int)R.get(3); val = 0; }
            else if (sym == 18) { rep = 11 + (int)R.get(7); val = 0; }
            if (i + rep > total) return false;
            // lens[] index: lit/len symbols at 0.., distance symbols at 288..
            for (int r = lane; r < rep; r += 32) {
                int idx = i + r;
                S.lens[idx < nlit ? idx : 288 + (idx - nlit)] = (uint8_t)val;
            }
            i += rep;
            prev = val;
        }
        __syncwarp();
        if (S.lens[256] == 0) return false;
    }
    return inf_build(S, 0, nlit, lane) && inf_build(S, 1, ndist, lane);

}

// Decodes one stream. All 32 lanes of the warp must call this with identical arguments.
__device__ inline void inflate_stream(InflateJob& job, InflateSmem& S, int lane)
{
    InflateReader R;
    R.words = (const uint32_t*)job.in;
    R.nwords = (job.in_len + 8 + 3) >> 2;
    R.lane = lane;
    uint8_t* out = job.out;
    const uint32_t cap = job.out_cap;
    const uint32_t in_len = job.in_len;
    uint32_t pos = 0;
    int status = INF_OK;
    uint32_t start = 0;

    if (job.parse_header) {
        if (in_len < 2) { status = INF_DATA_ERROR; goto done; }
        uint32_t cmf = job.in[0], flg = job.in[1];
        if (((cmf * 256 + flg) % 31 != 0) || (flg & 32) || ((cmf & 15) != 8)) { status = INF_DATA_ERROR; goto done; }
        start = 2;
    }
    R.seek(start);

    for (;;) {
        R.refill();
        uint32_t bfinal = R.get(1);
        uint32_t btype = R.get(2);
        if (btype == 0) {
            // stored block
            R.drop(R.bitcnt & 7);
            R.refill();
            uint32_t len = R.get(16);
            R.refill();
            uint32_t nlen = R.get(16);
            if ((len ^ 0xffffu) != nlen) { status = INF_DATA_ERROR; goto done; }
            uint32_t bp = R.bytepos_ceil();
            if (bp + len > in_len) { status = INF_DATA_ERROR; goto done; }
            uint32_t n = len;
            if (pos + n > cap) { n = cap - pos; status = INF_OUTPUT_FULL; }
            for (uint32_t i = lane; i < n; i += 32) out[pos + i] = job.in[bp + i];
            pos += n;
            if (status) goto done;
            __syncwarp();
            R.seek(bp + len);
        } else if (btype == 3) {
            status = INF_DATA_ERROR; goto done;
        } else {
            if (!inf_setup_tables(R, S, lane, btype)) { status = INF_DATA_ERROR; goto done; }

            // ---- symbol loop ----
            int nl = 0;             // literals batched in registers
            uint32_t mylit = 0;
            for (;;) {
                R.refill();
                uint32_t e = S.lit_tab[R.peek(INF_LIT_BITS)];
                if ((e & 15) == 0) {
                    e = inf_slow(S, 0, R.peek(15), INF_LIT_BITS);
                    if (e == 0) { status = INF_DATA_ERROR; break; }
                }
                R.drop(e & 15);
                uint32_t kind = (e >> 8) & 3;
                if (kind == 0) {
                    if (lane == nl) mylit = e >> 16;
                    if (++nl == 32) {
                        if (pos + 32 > cap) { status = INF_OUTPUT_FULL; break; }
                        out[pos + lane] = (uint8_t)mylit;
                        pos += 32; nl = 0;
                    }
                    continue;
                }
                // flush pending literals before anything that reads or ends the output
                if (nl) {
                    if (pos + nl > cap) { status = INF_OUTPUT_FULL; break; }
                    if (lane < nl) out[pos + lane] = (uint8_t)mylit;
                    pos += nl; nl = 0;
                }
                if (kind == 2) break;                       // end of block
                if (kind == 3) { status = INF_DATA_ERROR; break; }
                uint32_t xb = (e >> 4) & 15;
                uint32_t len = (e >> 16) + R.get(xb);
                R.refill();
                uint32_t de = S.dist_tab[R.peek(INF_DIST_BITS)];
                if ((de & 15) == 0) {
                    de = inf_slow(S, 1, R.peek(15), INF_DIST_BITS);
                    if (de == 0) { status = INF_DATA_ERROR; break; }
                }
                R.drop(de & 15);
                if (((de >> 8) & 3) == 3) { status = INF_DATA_ERROR; break; }
                R.refill();
                uint32_t dist = (de >> 16) + R.get((de >> 4) & 15);
                if (dist > pos) { status = INF_DATA_ERROR; break; }
                uint32_t n = len;
                if (pos + n > cap) { n = cap - pos; status = INF_OUTPUT_FULL; }
                __syncwarp();
                if (dist >= 32) {
                    for (uint32_t b = 0; b < n; b += 32) {
                        uint32_t i = b + lane;
                        if (i < n) out[pos + i] = out[pos + i - dist];
                        __syncwarp();
                    }
                } else {
                    // overlapping match: every byte comes from the `dist` bytes before pos
                    for (uint32_t b = 0; b < n; b += 32) {
                        uint32_t i = b + lane;
                        if (i < n) out[pos + i] = out[pos - dist + (i % dist)];
                    }
                    __syncwarp();
                }
                pos += n;
                if (status) break;
                // input exhausted long ago? (garbage guard)
                if (R.widx * 4 > in_len + 64) { status = INF_DATA_ERROR; break; }
            }
            if (status) goto done;
        }
        if (bfinal) break;
        if (R.widx * 4 > in_len + 64) { status = INF_DATA_ERROR; goto done; }
    }
    // all consumed bits must lie inside the input (miniz: running out of input is a data error)
    if (status == INF_OK) {
        uint64_t bits_used = (uint64_t)R.widx * 32 - (uint64_t)R.bitcnt;
        if (bits_used > (uint64_t)in_len * 8) status = INF_DATA_ERROR;
    }
done:
    __syncwarp();
    if (lane == 0) { job.out_len = pos; job.status = status; }
}

} // namespace gb
