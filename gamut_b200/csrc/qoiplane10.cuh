// qoiplane10.cuh -- parallel QOI-Plane10 decode (included by qoix.cu).
//
// qoiplane10_decode (codecs/qoiplane10.d:317-515) is two serial dependencies wound together: the opcode
// stream (2-bit aligned variable-length codes, every position depends on all codes before it) and the
// pixel recurrence (l = MED(left, up, upleft) + residual). They are taken apart:
//
//  A. parse, chunk-parallel.  Opcode lengths depend on the bits only, never on pixel values, and a parser
//     started at a wrong position falls onto true code boundaries after a few codes. The stream is cut into
//     128-byte chunks, one thread each:
//       p10_sync_kernel   a CTA takes 248 consecutive chunks (+ 8 warm-up chunks of its predecessor) with its slice
//                         of the stream in shared memory. Every thread parses its chunk (after a 96-bit uncounted
//                         run-up), then the CTA relaxes: chunks whose predecessor's exit differs from the entry
//                         they used parse again, until nothing changes. Chunk 0 starts at the true position, so
//                         the fixed point is the serial parse (induction over chunks). Every parse also records
//                         the chunk's pixel count, its effect on the running alpha (set v / add d) and whether it
//                         met END. p10_repair_kernel checks the entry of every CTA's first own chunk.
//       p10_scan_kernel   one CTA per image: exclusive scan of (pixels, alpha transform, ended) over chunks.
//       p10_write_kernel  parses once more from the true entry state and writes one 32-bit record per pixel:
//                         residual, COPY (of the previous pixel) / LIT flags, the pixel's alpha.
//  B. p10_recon_kernel, eight warps per image pipelined over bands of 32 rows, each band a skewed wavefront: lane r runs one 8-pixel
//     block behind lane r-1, so the row above arrives by warp shuffle off the critical path and the chain
//     per pixel is just the MED predictor. A run that crosses a row boundary makes the first pixel of a row
//     depend on the end of the row above; such rows start later (exactly as late as the dependency demands)
//     and read the row above from memory.
#pragma once

constexpr int P10_CHUNK_BYTES = 128;
constexpr int P10_CHUNK_BITS = P10_CHUNK_BYTES * 8;
constexpr int P10_CHUNK_WORDS = P10_CHUNK_BYTES / 4;
// per-pixel record: bits 0-9 residual (LIT: the value itself), bit 10 COPY of the previous pixel, bit 11 LIT,
// bits 16-31 the pixel's alpha already expanded to 16 bits
constexpr uint32_t P10_REC_COPY = 1u << 10, P10_REC_LIT = 1u << 11;
constexpr int P10_CTA = 256;                     // threads (= chunk slots) per CTA of the sync and write kernels
constexpr uint32_t P10_RUNUP_BITS = 96;           // sync kernel: the first pass starts this far before its chunk
constexpr int P10_WARM = 8;                      // sync: slots that re-parse the tail of the previous CTA's range
constexpr int P10_OWN = P10_CTA - P10_WARM;
constexpr int P10_STAGE_WORDS = P10_CTA * P10_CHUNK_WORDS + 32;     // the CTA's slice of the stream + what a parse may read past it

struct P10Image {
    const uint8_t* stream;      // payload with its 25-byte header
    uint32_t size;              // bytes in `stream`
    uint32_t w, h, wp;          // wp = record pitch (w rounded up to 8)
    int channels;
    int image;                  // index into status[]
    uint32_t chunk_base, nchunks;
    uint32_t scta_base, wcta_base;      // first CTA of this image in the sync grid (P10_OWN chunks each) / write grid (P10_CTA)
    uint32_t* recs;             // [h][wp]
    uint32_t* rowinfo;          // [h]: 0 = row starts normally, x+2 = first pixel copies column x of the row above (x = -1: any)
    uint8_t* out;
};

struct P10Chunk {               // per-chunk parse summary
    uint32_t exit_bit;          // position of the first opcode that starts at or after the chunk's end
    uint32_t npix;              // pixels produced by the opcodes that start in this chunk
    uint32_t alpha;             // bit 31: ended, bit 30: set, bits 0-9: value (set) or delta (add)
};
struct P10Entry { uint32_t pix; uint32_t alpha; };     // alpha bit 31: the stream ended before this chunk

template <int WHICH>            // 0: chunk_base, 1: scta_base, 2: wcta_base
__device__ __forceinline__ uint32_t p10_find_image(const P10Image* __restrict__ imgs, int n, uint32_t c)
{
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        const uint32_t b = WHICH == 0 ? imgs[mid].chunk_base : WHICH == 1 ? imgs[mid].scta_base : imgs[mid].wcta_base;
        if (b <= c) lo = mid; else hi = mid - 1;
    }
    return (uint32_t)lo;
}

// The opcode stream as big-endian 32-bit words: word k = payload bits [32k, 32k + 32). Bytes past the end read as
// 0xFF (== END), like the restated bit reader of the oracle.
struct P10Global {
    const uint32_t* words; uint32_t adj, nbytes;
    __device__ __forceinline__ void init(const uint8_t* stream, uint32_t size)
    {
        const uintptr_t a = (uintptr_t)(stream + 25);
        words = (const uint32_t*)(a & ~(uintptr_t)3); adj = (uint32_t)(a & 3) * 8;
        nbytes = size - 25;
    }
    __device__ __forceinline__ uint32_t operator()(long long k) const
    {
        if (k < 0) return 0xffffffffu;
        const unsigned long long b = (unsigned long long)k * 4;
        if (b + 8 <= nbytes) {
            const uint32_t w0 = __byte_perm(__ldg(words + k), 0, 0x0123), w1 = __byte_perm(__ldg(words + k + 1), 0, 0x0123);
            return __funnelshift_l(w1, w0, adj);
        }
        const uint8_t* p = (const uint8_t*)words + (adj >> 3);
        uint32_t v = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) v = (v << 8) | (b + i < nbytes ? (uint32_t)p[b + i] : 0xFFu);
        return v;
    }
};
// The CTA's slice of the stream in shared memory: word k of the slice at [k / 32][(k + k / 32) % 32], so that the
// lanes of a warp, each reading its own 128-byte chunk, hit different banks.
struct P10Shared {
    const uint32_t* s;
    __device__ __forceinline__ static uint32_t swz(uint32_t k) { return (k & ~31u) | ((k + (k >> 5)) & 31u); }
    __device__ __forceinline__ uint32_t operator()(uint32_t k) const { return s[swz(k)]; }
};
__device__ __forceinline__ void p10_stage(uint32_t* s_words, const P10Global& G, long long first_word, int tid)
{
    for (int k = tid; k < P10_STAGE_WORDS; k += P10_CTA) s_words[P10Shared::swz((uint32_t)k)] = G(first_word + k);
}

// One table lookup per opcode instead of a chain of data-dependent branches (the lanes of a warp are at different
// opcodes): length, pixels produced, and where the value sits in the 32 bits at the opcode's start.
//   bits 0-4 length | 5-8 pixels | 9 extended run | 10 alpha opcode (ADIFF; with pixels = 1: LA) | 11 END / reserved
//   | 12-16, 17-21 the left / arithmetic right shift that extract the sign-extended value | 20-21 (<< 2) kind
constexpr uint32_t P10L_EXT = 1u << 9, P10L_ALPHA = 1u << 10, P10L_END = 1u << 11, P10L_NPIX = 15u << 5;
__device__ __forceinline__ uint32_t p10_lut_entry(uint32_t op)      // opcode table, qoiplane10.d:43-52
{
    auto val = [](int shift, int bits) { return (uint32_t)(32 - shift - bits) << 12 | (uint32_t)(32 - bits) << 17; };
    if (op < 0x80) return 4u | 1u << 5 | val(28, 3);                                   // DIFF1
    if (op < 0xc0) return 8u | 1u << 5 | val(24, 6);                                   // DIFF2
    if (op < 0xe0) {                                                                   // RUN
        const uint32_t run = (op >> 2) & 7;
        return run == 7 ? (14u | P10L_EXT | 1u << 22) : (6u | (run + 1) << 5 | 1u << 22);
    }
    if (op < 0xf0) return 14u | 1u << 5 | val(18, 10);                                 // DIFF4
    if (op < 0xf8) return 12u | 1u << 5 | val(20, 7);                                  // DIFF3
    if (op < 0xfc) return 12u | P10L_ALPHA;                                            // ADIFF (the pixel's own opcode follows)
    if (op == 0xfe) return 28u | 1u << 5 | P10L_ALPHA | val(14, 10) | 2u << 22;        // LA
    return P10L_END;                                                                   // END (0xff), reserved 0xfc / 0xfd
}

// Bit reader over a word accessor, MSB first: two words of the stream and the bit offset into the first. The refill
// is straight-line code (the next word is always loaded, selects decide whether it is used): the lanes of a warp
// cross their word boundaries at different opcodes, and a branch there would run for a handful of lanes at a time.
template <class Words>
struct P10Reader {
    Words W; uint32_t next, hi, lo, sh;
    __device__ __forceinline__ void init(uint32_t bitpos)
    {
        next = bitpos >> 5;
        hi = W(next); lo = W(next + 1);
        next += 2;
        sh = bitpos & 31;
    }
    __device__ __forceinline__ uint32_t peek() const { return __funnelshift_l(lo, hi, sh); }
    __device__ __forceinline__ void drop(int n)
    {
        sh += (uint32_t)n;
        const uint32_t nw = W(next);
        const bool adv = sh >= 32;
        hi = adv ? lo : hi; lo = adv ? nw : lo;
        next += adv ? 1u : 0u; sh &= 31u;
    }
};

// Parses the opcode groups that START in [bitpos, limit) without storing anything (positions are relative to the word
// accessor's origin): exit position, pixels produced, effect on the running alpha. A group is a pixel's opcode with
// the ADIFFs before it; an ADIFF adds to the alpha of the PREVIOUS pixel (qoiplane10.d:463-470), so of several
// ADIFFs in one group only the last counts, and a group belongs to the chunk it starts in.
template <class Words>
__device__ __forceinline__ bool p10_count(const Words& W, const uint32_t* __restrict__ lut, uint32_t& bitpos, uint32_t limit,
                                          uint32_t& npix, uint32_t& alpha_fx)
{
    P10Reader<Words> R{W}; R.init(bitpos);
    uint32_t p = bitpos, np = 0;
    bool aset = false, ended = false, ing = false; int aval = 0, gd = 0;
    while (p < limit || ing) {
        const uint32_t v = R.peek();
        const uint32_t e = lut[v >> 24];
        const uint32_t len = e & 31u;
        if ((e & (P10L_ALPHA | P10L_END)) || ing) {              // rare: the alpha plane of the workload is mostly flat
            if (e & P10L_END) { ended = true; break; }
            if ((e & (P10L_ALPHA | P10L_NPIX)) == P10L_ALPHA) { gd = (int)(v << 6) >> 26; ing = true; p += len; R.drop((int)len); continue; }
            if (e & P10L_ALPHA) { aset = true; aval = (int)((v >> 4) & 1023u); }
            else aval += gd;
            ing = false;
        }
        np += (e & P10L_EXT) ? ((v >> 18) & 0xffu) + 8u : (e >> 5) & 15u;
        p += len; R.drop((int)len);
    }
    bitpos = p; npix = np;
    alpha_fx = (aset ? 0x40000000u : 0u) | ((uint32_t)aval & 1023u);
    return ended;
}

// ---- A1. speculative parse + CTA-local relaxation ----------------------------------------------------------------
// One CTA = 248 consecutive chunks of one image plus 8 warm-up chunks of its predecessor, the slice of the stream in
// shared memory. Every thread parses its chunk from its first bit, then the CTA relaxes: chunks whose predecessor's
// exit differs from the entry they used are compacted into as few warps as possible and parse again, until nothing
// changes. Chunk 0 starts at the true position, so the fixed point is the serial parse. The one unverified
// assumption, the entry of the CTA's first own chunk, is checked by p10_repair_kernel.
__global__ void __launch_bounds__(P10_CTA)
p10_sync_kernel(const P10Image* __restrict__ imgs, int nimgs, P10Chunk* __restrict__ chunks, uint32_t* __restrict__ entry_used)
{
    __shared__ uint32_t s_words[P10_STAGE_WORDS];
    __shared__ uint32_t s_lut[256];
    __shared__ uint32_t s_entry[P10_CTA], s_exit[P10_CTA], s_npix[P10_CTA], s_alpha[P10_CTA], s_todo_entry[P10_CTA];
    __shared__ uint16_t s_todo[P10_CTA];
    __shared__ uint32_t s_wcount[P10_CTA / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const P10Image& im = imgs[p10_find_image<1>(imgs, nimgs, blockIdx.x)];
    const uint32_t local_cta = blockIdx.x - im.scta_base;
    const int lc_first = (int)(local_cta * P10_OWN) - P10_WARM;            // chunk of slot 0 (negative in the first CTA)
    P10Global G; G.init(im.stream, im.size);
    const uint32_t total_bits = G.nbytes * 8;
    s_lut[tid] = p10_lut_entry((uint32_t)tid);
    p10_stage(s_words, G, (long long)lc_first * P10_CHUNK_WORDS, tid);
    const P10Shared W{s_words};
    const int origin = lc_first * P10_CHUNK_BITS;                           // stream position of bit 0 of the slice
    const int lc = lc_first + tid;
    const bool active = lc >= 0 && (uint32_t)lc < im.nchunks;
    const uint32_t limit = min((uint32_t)(lc + 1) * (uint32_t)P10_CHUNK_BITS, total_bits);
    auto parse = [&](uint32_t entry, uint32_t& exit_bit, uint32_t& npix, uint32_t& alpha) {
        uint32_t rel = entry - (uint32_t)origin;
        const bool ended = p10_count(W, s_lut, rel, limit - (uint32_t)origin, npix, alpha) || rel + (uint32_t)origin >= total_bits;
        // A parse that stopped at END leaves a neutral exit (the next chunk's own first bit): after a true END the later
        // chunks do not matter, and after a false one (a speculative parse on garbage) the successor keeps its own guess.
        exit_bit = ended ? (uint32_t)(lc + 1) * (uint32_t)P10_CHUNK_BITS : rel + (uint32_t)origin;
        if (ended) alpha |= 0x80000000u;
    };
    __syncthreads();
    {
        // A parse started on the chunk's first bit is wrong more often than not (an opcode straddles the boundary),
        // but opcodes fall into step within a few codes: the parse starts P10_RUNUP_BITS early, uncounted, and has
        // usually found the true boundary by the time it reaches its chunk.
        uint32_t entry = active ? (uint32_t)lc * (uint32_t)P10_CHUNK_BITS : 0u;
        if (active && lc > 0 && tid > 0) {                          // (slot 0's run-up would lie before the staged slice)
            uint32_t rel = entry - P10_RUNUP_BITS - (uint32_t)origin, np0 = 0, al0 = 0;
            const bool e0 = p10_count(W, s_lut, rel, entry - (uint32_t)origin, np0, al0);
            if (!e0) entry = rel + (uint32_t)origin;                // ended (on garbage): keep the chunk's own first bit
        }
        uint32_t x = total_bits, np = 0, al = 0;
        if (active) parse(entry, x, np, al);
        s_entry[tid] = entry; s_exit[tid] = x; s_npix[tid] = np; s_alpha[tid] = al;
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
        for (int w = 0; w < P10_CTA / 32; ++w) { const uint32_t c = s_wcount[w]; off += w < warp ? c : 0; total += c; }
        if (total == 0) break;
        if (stale) { const uint32_t q = off + __popc(bal & ((1u << lane) - 1)); s_todo[q] = (uint16_t)tid; s_todo_entry[q] = prev; }
        __syncthreads();
        if ((uint32_t)tid < total) {
            const int s = s_todo[tid];
            const uint32_t entry = s_todo_entry[tid];
            const int slc = lc_first + s;
            uint32_t rel = entry - (uint32_t)origin, np = 0, al = 0;
            const uint32_t lim = min((uint32_t)(slc + 1) * (uint32_t)P10_CHUNK_BITS, total_bits);
            const bool ended = p10_count(W, s_lut, rel, lim - (uint32_t)origin, np, al) || rel + (uint32_t)origin >= total_bits;
            s_entry[s] = entry;
            s_exit[s] = ended ? (uint32_t)(slc + 1) * (uint32_t)P10_CHUNK_BITS : rel + (uint32_t)origin;
            s_npix[s] = np; s_alpha[s] = al | (ended ? 0x80000000u : 0u);
        }
    }
    if (tid >= P10_WARM && active) {
        P10Chunk* dst = chunks + im.chunk_base + (uint32_t)lc;
        dst->exit_bit = s_exit[tid]; dst->npix = s_npix[tid]; dst->alpha = s_alpha[tid];
    }
    if (tid == P10_WARM) entry_used[blockIdx.x] = s_entry[tid];
}

// ---- A1b. repair / check of the CTA boundaries --------------------------------------------------------------------
// One thread per sync CTA compares the entry its first own chunk used with the true exit of the chunk before it and,
// where they differ (rare), parses forward through its own range until it meets the old chain. mode 1 only counts
// the boundaries that still disagree; the host reads that count with the statuses at the very end.
__global__ void __launch_bounds__(64)
p10_repair_kernel(const P10Image* __restrict__ imgs, int nimgs, uint32_t total_ctas, P10Chunk* chunks, uint32_t* entry_used,
                  int mode, uint32_t* unconverged)
{
    __shared__ uint32_t s_lut[256];
    for (int i = threadIdx.x; i < 256; i += 64) s_lut[i] = p10_lut_entry((uint32_t)i);
    __syncthreads();
    const uint32_t cta = blockIdx.x * blockDim.x + threadIdx.x;
    if (cta >= total_ctas) return;
    const P10Image& im = imgs[p10_find_image<1>(imgs, nimgs, cta)];
    const uint32_t local_cta = cta - im.scta_base;
    if (local_cta == 0) return;                                  // starts from the true position
    const uint32_t lc0 = local_cta * P10_OWN;
    if (lc0 >= im.nchunks) return;
    P10Chunk* const ch = chunks + im.chunk_base;
    const uint32_t truth = __ldcg(&ch[lc0 - 1].exit_bit);
    if (truth == entry_used[cta]) return;
    if (mode == 1) { atomicAdd(unconverged, 1u); return; }
    entry_used[cta] = truth;
    P10Global G; G.init(im.stream, im.size);
    const uint32_t total_bits = G.nbytes * 8;
    const uint32_t lc_end = min(lc0 + (uint32_t)P10_OWN, im.nchunks);
    uint32_t bp = truth;
    for (uint32_t lc = lc0; lc < lc_end; ++lc) {
        uint32_t np = 0, al = 0;
        const bool ended = p10_count(G, s_lut, bp, min((lc + 1) * (uint32_t)P10_CHUNK_BITS, total_bits), np, al) || bp >= total_bits;
        if (ended) { bp = (lc + 1) * (uint32_t)P10_CHUNK_BITS; al |= 0x80000000u; }
        const uint32_t old = __ldcg(&ch[lc].exit_bit);
        __stcg(&ch[lc].exit_bit, bp); ch[lc].npix = np; ch[lc].alpha = al;
        if (old == bp) break;                                    // met the old chain: everything after is already right
    }
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
// Every chunk is parsed once more from its true entry state. A lane collects its records four at a time in
// registers and stores them as one aligned 16-byte vector (single records only at the two ends of its part of the
// image); long runs are placed by the whole warp, 32 records per step, straight into memory.
__global__ void __launch_bounds__(P10_CTA)
p10_write_kernel(const P10Image* __restrict__ imgs, int nimgs, const P10Chunk* __restrict__ chunks,
                 const P10Entry* __restrict__ entries)
{
    __shared__ uint32_t s_words[P10_STAGE_WORDS];
    __shared__ uint32_t s_lut[256];
    const int tid = threadIdx.x, lane = tid & 31;
    const P10Image& im = imgs[p10_find_image<2>(imgs, nimgs, blockIdx.x)];
    const uint32_t lc0 = (blockIdx.x - im.wcta_base) * P10_CTA;
    P10Global G; G.init(im.stream, im.size);
    const uint32_t total_bits = G.nbytes * 8;
    s_lut[tid] = p10_lut_entry((uint32_t)tid);
    p10_stage(s_words, G, (long long)lc0 * P10_CHUNK_WORDS, tid);
    const P10Shared W{s_words};
    const uint32_t origin = lc0 * (uint32_t)P10_CHUNK_BITS;
    const uint32_t lc = lc0 + tid;
    const uint32_t np = im.w * im.h;                              // < 400e6 (header check)
    bool alive = lc < im.nchunks;
    uint32_t i = 0, a = 0, bp = 0;
    if (alive) {
        const P10Entry e = entries[im.chunk_base + lc];
        i = e.pix; a = e.alpha & 1023u;
        bp = lc ? chunks[im.chunk_base + lc - 1].exit_bit : 0u;
        alive = !(e.alpha & 0x80000000u) && i < np;
    }
    const uint32_t limit = min((lc + 1) * (uint32_t)P10_CHUNK_BITS, total_bits);
    alive = alive && bp < limit;
    const uint32_t Wd = im.w, WP = im.wp, skip = im.wp - im.w;
    uint32_t y = alive ? i / Wd : 0u, x = alive ? i - y * Wd : 0u;
    uint32_t addr = y * WP + x;                                   // index into the record array
    uint32_t q0 = 0, q1 = 0, q2 = 0, q3 = 0, qmask = 0;           // records of the aligned group of four `addr` is in
    uint32_t a_pend = 0; bool ing = false;                        // inside a group: alpha an ADIFF has announced
    uint32_t* const recs = im.recs; uint32_t* const rowinfo = im.rowinfo;
    P10Reader<P10Shared> R{W};
    __syncthreads();
    R.init(alive ? bp - origin : 0u);
    // what the first pixel of row yy records when it is pixel ii of a COPY run that started at pixel p0
    auto row_info = [&](uint32_t rec, uint32_t ii, uint32_t yy, uint32_t p0) {
        uint32_t info = 0;
        if (yy > 0 && (rec & P10_REC_COPY)) {
            // the pixel copies pixel p0-1 (and everything between is a copy of it too)
            const uint32_t dist = ii - (p0 - 1);
            info = dist <= Wd ? (Wd - dist) + 2u : 1u;
        }
        return info;
    };
    auto flush_group = [&]() {           // the group that ends just before (or contains) addr - 1
        uint32_t* g = recs + ((addr - 1) & ~3u);
        if (qmask == 15u) *(uint4*)g = make_uint4(q0, q1, q2, q3);
        else {
            if (qmask & 1u) g[0] = q0;
            if (qmask & 2u) g[1] = q1;
            if (qmask & 4u) g[2] = q2;
            if (qmask & 8u) g[3] = q3;
        }
        qmask = 0;
    };
    // One straight-line round per opcode: a lane decodes when it has nothing left to place, then places ONE record
    // (a short run takes several rounds; its lane does not decode meanwhile). The kinds of opcode differ in a few
    // selects, not in the path taken: the lanes of a warp are at different opcodes.
    uint32_t pend_n = 0, pend_rec = 0, pend_p0 = 0;
    while (__any_sync(0xffffffffu, alive || pend_n)) {
        const bool dec = alive && pend_n == 0;
        const uint32_t v = R.peek();
        const uint32_t e = s_lut[v >> 24];
        const bool end = dec && (e & P10L_END);
        const bool adiff = dec && (e & (P10L_ALPHA | P10L_NPIX | P10L_END)) == P10L_ALPHA;     // the pixel's own opcode follows
        const bool pix = dec && !end && !adiff;
        const uint32_t len = (dec && !end) ? e & 31u : 0u;
        a_pend = adiff ? (a + (uint32_t)((int)(v << 6) >> 26)) & 1023u : a_pend;
        a = pix ? ((e & P10L_ALPHA) ? (v >> 4) & 1023u : (ing ? a_pend : a)) : a;             // LA sets, ADIFF announced
        ing = adiff || (ing && !pix);
        {
            uint32_t n = (e & P10L_EXT) ? ((v >> 18) & 0xffu) + 8u : (e >> 5) & 15u;
            n = pix ? min(n, np - i) : 0u;
            const uint32_t kind = ((e >> 22) & 3u) << 10;
            const uint32_t val = (kind & P10_REC_COPY) ? 0u : (uint32_t)((int)(v << ((e >> 12) & 31u)) >> ((e >> 17) & 31u)) & 1023u;
            pend_rec = dec ? (kind | val | (((a << 6) | (a >> 4)) << 16)) : pend_rec;
            pend_n = dec ? n : pend_n;
            pend_p0 = dec ? i : pend_p0;
        }
        bp += len; R.drop((int)len);
        alive = alive && !end && !(pix && bp >= limit);           // the next group starts in a later chunk
        // long runs: the whole warp places the records of one lane's run straight into memory, 32 per step (a lane
        // on its own would keep the other 31 waiting for up to 262 steps)
        uint32_t big = __ballot_sync(0xffffffffu, pend_n >= 8);
        if (big) {
            if (pend_n >= 8 && qmask) flush_group();
            while (big) {
                const int l = __ffs(big) - 1; big &= big - 1;
                const uint32_t rn = __shfl_sync(0xffffffffu, pend_n, l), rrec = __shfl_sync(0xffffffffu, pend_rec, l);
                const uint32_t ri = __shfl_sync(0xffffffffu, i, l), rx = __shfl_sync(0xffffffffu, x, l), ry = __shfl_sync(0xffffffffu, y, l);
                for (uint32_t j = lane; j < rn; j += 32) {
                    const uint32_t xx0 = rx + j, dy = xx0 / Wd, xx = xx0 - dy * Wd, yy = ry + dy;
                    recs[(size_t)yy * WP + xx] = rrec;
                    if (xx == 0) rowinfo[yy] = row_info(rrec, ri + j, yy, ri);
                }
                if (lane == l) {
                    const uint32_t xx0 = x + pend_n, dy = xx0 / Wd;
                    i += pend_n; x = xx0 - dy * Wd; y += dy; addr = y * WP + x;
                    pend_n = 0;
                }
            }
        }
        // one record
        const bool place = pend_n != 0;
        if (place && x == 0) rowinfo[y] = row_info(pend_rec, i, y, pend_p0);
        const uint32_t k = addr & 3u;
        q0 = (place && k == 0) ? pend_rec : q0; q1 = (place && k == 1) ? pend_rec : q1;
        q2 = (place && k == 2) ? pend_rec : q2; q3 = (place && k == 3) ? pend_rec : q3;
        qmask |= place ? 1u << k : 0u;
        const uint32_t adv = place ? 1u : 0u;
        pend_n -= adv; i += adv; addr += adv; x += adv;
        if (place && k == 3) {
            if (qmask == 15u) { *(uint4*)(recs + (addr - 4)) = make_uint4(q0, q1, q2, q3); qmask = 0; }
            else flush_group();
        }
        if (x == Wd) { x = 0; ++y; if (skip) { if (qmask) flush_group(); addr += skip; } }
        if (i >= np) { alive = false; pend_n = 0; }
        if (!alive && pend_n == 0 && qmask) flush_group();
    }
}

__device__ __forceinline__ int p10_sext(uint32_t v, int bits) { return (int)(v << (32 - bits)) >> (32 - bits); }

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
        // l of 8 pixels of the finished row above, from memory (written by another warp or another lane: L2 loads)
        auto load_up = [&](uint32_t xx0, int* dst) {
            if (vec_ok) {
                if (CH == 2) {
                    const uint4 a = __ldcg((const uint4*)(out16 + ((size_t)(y - 1) * W + xx0) * 2)), b = __ldcg((const uint4*)(out16 + ((size_t)(y - 1) * W + xx0) * 2) + 1);
                    const uint32_t wv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) dst[i] = (int)((wv[i] & 0xffffu) >> 6);
                } else {
                    const uint4 a = __ldcg((const uint4*)(out16 + (size_t)(y - 1) * W + xx0));
                    const uint32_t wv[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) dst[i] = (int)(((wv[i >> 1] >> ((i & 1) * 16)) & 0xffffu) >> 6);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) dst[i] = xx0 + i < W ? p10_out_l<CH>(im.out, W, y - 1, xx0 + i) : 0;
            }
        };

        for (uint32_t T = 0; T < total; ++T) {
            const bool active = row_ok && T >= start && T < start + nblocks;
            const uint32_t blk = T - start, x0 = blk << 3;
            int up[8];
            // the row above: shuffle the previous block results of the lane above, or memory
#pragma unroll
            for (int i = 0; i < 8; ++i) up[i] = __shfl_up_sync(0xffffffffu, prevres[i], 1);
            // memory phase. The wait for the warp that owns the band above is taken by the whole warp on lane 0's
            // behalf (one uniform loop instead of one lane leaving the others behind).
            uint4 na = ra, nb = rb;
            if (y0 > 0) {
                uint32_t need = 0;
                if (lane == 0 && active) {
                    // the row above belongs to another warp: wait until it is in memory as far as this step reads it
                    need = min(blk + 2, nblocks);
                    if (blk == 0 && wrap) need = max(need, min((uint32_t)(max(xn, 0) >> 3) + 1u, nblocks));
                    need += kprev_base;
                }
                need = __shfl_sync(0xffffffffu, need, 0);
                while (prog[pw] < need) __nanosleep(20);
            }
            if (active) {
                if (blk == 0) {
                    ra = __ldg((const uint4*)rrow); rb = __ldg((const uint4*)rrow + 1);
                    if (frommem) {
                        load_up(0, upm);
                        if (wrap) left = p10_out_l<CH>(im.out, W, y - 1, (uint32_t)max(xn, 0));
                    }
                }
                if (frommem) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) up[i] = upm[i];
                }
                // prefetch the next block
                if (blk + 1 < nblocks) {
                    na = __ldg((const uint4*)(rrow + x0 + 8)); nb = __ldg((const uint4*)(rrow + x0 + 12));
                    if (frommem) load_up(x0 + 8, upm);
                }
            }
            __syncwarp();
            // arithmetic phase: every lane, converged (lanes outside their row work on stale registers; nothing of
            // theirs is stored, and a lane's registers are set up again at its block 0)
            {
                if (y == 0) {                       // row 0 predicts from the left neighbour alone: median(left, 0, left + 0 - 0)
#pragma unroll
                    for (int i = 0; i < 8; ++i) up[i] = 0;
                }
                // The first pixel of a row is predicted by the pixel above it (a COPY keeps the previous pixel in
                // raster order, which `left` holds when the row starts inside a run): median(up, up, up).
                if (blk == 0) { upleft = up[0]; left = (ra.x & P10_REC_COPY) ? left : up[0]; }
                const uint32_t recs[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
                uint32_t o[8];
                // locoPredict (qoiplane10.d:84-96) is the median of left, top and left + top - topleft; one straight-line
                // body per pixel: a COPY record zeroes top - topleft (the median of left, top, left is left) and
                // carries residual 0, a LIT record masks the prediction away.
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t r = recs[i];
                    const uint32_t keep = (((r >> 11) & 1u) - 1u) & 1023u;       // 0 for LIT, else 1023
                    const uint32_t res = r & 1023u;
                    const int d = (r & P10_REC_COPY) ? 0 : up[i] - (i ? up[i - 1] : upleft);
                    const int t = left + d, mn = min(left, up[i]), mx = max(left, up[i]);
                    const int pred = max(mn, min(mx, t));
                    const int l = (int)((((uint32_t)pred + (res & keep)) & keep) | (res & ~keep));
                    left = l;
                    prevres[i] = l;
                    const uint32_t l16 = (uint32_t)((l << 6) | (l >> 4));
                    o[i] = CH == 2 ? __byte_perm(l16, r, 0x7610) : l16;
                }
                upleft = up[7];
                const unsigned long long gi0 = (unsigned long long)y * W + x0;
                if (active && gi0 + 8 > ndec) {     // the stream ended inside or before this block: the rest stays zero
#pragma unroll
                    for (int i = 0; i < 8; ++i) if (gi0 + i >= ndec) { o[i] = 0; prevres[i] = 0; }
                    left = prevres[7];
                }
                if (active) {
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
            }
            __syncwarp();       // orders this block's stores before the loads of the lanes that read rows from memory
        }
        __syncwarp();
        if (lane == 0) { __threadfence_block(); prog[warp] = (kband + 1) * (nblocks + 1); }
    }
}
