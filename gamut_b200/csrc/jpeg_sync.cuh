// jpeg_sync.cuh -- intra-image parallel Huffman decoding for long entropy segments (included by jpeg.cu).
//
// A baseline JPEG scan without restart markers is one serial bit stream (decode_next_row,
// jpegload.d:2405-2525). Huffman codes self-synchronise, so the stream is cut into 512-byte chunks, one
// thread each, and the whole stage runs without a host round trip:
//   1. jpeg_unstuff_*       (three small kernels over 4 KB tiles) remove the FF00 byte stuffing and stop at the first
//                           marker, giving a plain bit stream padded with 1-bits (the reference reads all-ones past a marker,
//                           jpegload.d:683-743);
//   2. jpeg_sync_kernel     one CTA = 248 consecutive chunks of one segment (+ 8 warm-up chunks borrowed from
//                           its predecessor). Every thread decodes its chunk from a guessed state (after a run-up of
//                           256 bytes, so that the guess is usually right by then), then the CTA
//                           relaxes in shared memory: a thread whose predecessor's exit state differs from the
//                           entry state it used decodes again, until nothing changes (chunk 0 of a segment starts
//                           from the true state, so the fixed point is the serial decode, by induction). Per chunk
//                           it leaves the exit state, the number of blocks completed and the sum of the DC
//                           differences per component;
//   3. jpeg_repair_kernel   the only unverified assumption is the state at each CTA's first own chunk (taken from
//                           its warm-up chunks). One thread per CTA boundary compares it with the true exit state
//                           of the previous CTA and, if they differ (rare), walks forward re-decoding chunks until
//                           the states meet again. jpeg_check_kernel counts boundaries that still disagree; the
//                           host reads that count with the statuses at the very end and only then repeats;
//   4. jpeg_scan_kernel     per segment: exclusive scans of the block counts and of the three DC sums;
//   5. jpeg_write_kernel    every chunk is decoded once more from its true state. A block belongs to the chunk
//                           its DC symbol starts in: that thread decodes it to its end (running into the next
//                           chunk if need be), zero-fills the 128 bytes, scatters the dequantised coefficients
//                           with the DC prediction already applied, and records the block's zig-zag extent.
#pragma once

constexpr int JS_CHUNK_BYTES = 512;
constexpr uint32_t JS_RUNUP_BITS = 2048;         // the first pass starts this far before its chunk (see jpeg_sync_kernel)
constexpr int JS_CHUNK_BITS = JS_CHUNK_BYTES * 8;
constexpr uint32_t JS_LONG_MIN = 1024;           // shorter segments are decoded by one thread each (jpeg_huffman_kernel)
constexpr int JS_CTA = 256;                      // threads (= chunks) per CTA of the sync kernel
constexpr int JS_WARM = 8;                       // of which warm-up chunks that belong to the previous CTA
constexpr int JS_OWN = JS_CTA - JS_WARM;
constexpr int JS_PAD_BYTES = 256;                // 1-bits after the unstuffed data (a block may run past the end)

struct LongSeg {
    int image;
    uint32_t in_start, in_end;      // raw (stuffed) byte range in the file
    uint32_t chunk_base, nchunks;   // this segment's slice of the per-chunk arrays (nchunks from the stuffed length)
    uint32_t cta_base;              // first sync CTA of this segment (JS_OWN chunks per CTA)
    int first_mcu, num_mcus;
    unsigned long long clean_off;   // offset of the unstuffed stream in the clean arena (16-byte aligned)
    uint32_t tile_base, ntiles;     // this segment's slice of the unstuff tiles (JU_TILE stuffed bytes each)
};

// decoder state between two symbols: bit position, block-in-MCU, zig-zag index (0 = DC next)
struct ChunkState { uint32_t bitpos; uint32_t bik; };               // bik = bi | k << 16
struct __align__(16) ChunkRec { uint32_t bitpos, bik, nblk; int dc0, dc1, dc2; uint32_t pad0, pad1; };   // 32 B
struct __align__(16) ChunkBase { uint32_t blk; int dc0, dc1, dc2; };                                     // 16 B

// The two-level tables of HuffTable (l1 then l2, contiguous) of the up to six tables of one image, in shared memory:
// table t = comp * 2 + (0 = DC, 1 = AC) at byte offset t * JS_TBL_BYTES. Addressed through 32-bit shared addresses
// (ld.shared with a register offset) rather than generic pointers.
constexpr uint32_t JS_L1_BYTES = sizeof(uint16_t) << HT_L1_BITS, JS_L2_BYTES = sizeof(uint16_t) * (HT_SUBS << HT_L2_BITS);
constexpr uint32_t JS_TBL_BYTES = JS_L1_BYTES + JS_L2_BYTES;
struct __align__(16) FastTables { uint8_t b[6 * JS_TBL_BYTES]; };
static_assert(offsetof(HuffTable, l2) == offsetof(HuffTable, l1) + JS_L1_BYTES, "l1 and l2 are copied as one block");

__device__ __forceinline__ uint32_t js_lds16(uint32_t saddr)
{
    uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(saddr)); return v;
}

__device__ __forceinline__ uint32_t js_find_seg(const LongSeg* __restrict__ segs, int nsegs, uint32_t cta)
{
    int lo = 0, hi = nsegs - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (segs[mid].cta_base <= cta) lo = mid; else hi = mid - 1; }
    return (uint32_t)lo;
}
__device__ __forceinline__ uint32_t js_find_seg_chunk(const LongSeg* __restrict__ segs, int nsegs, uint32_t c)
{
    int lo = 0, hi = nsegs - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (segs[mid].chunk_base <= c) lo = mid; else hi = mid - 1; }
    return (uint32_t)lo;
}

// ---- 1. unstuff: tiles of 4 KB of stuffed data, one CTA each ---------------------------------------------------
// Where an unstuffed byte lands depends on how many FF00 pairs precede it, and the data end at the first marker:
//   jpeg_unstuff_count  per tile: bytes kept (ignoring the marker) and the position of the first marker in the tile;
//   jpeg_unstuff_scan   per segment: first marker over the tiles, exclusive prefix of the counts up to its tile;
//   jpeg_unstuff_write  per tile: the kept bytes before the marker to their final positions; the tile that holds the
//                       end of the data also writes the length and the padding.
constexpr int JU_TILE = 4096;
struct JuTile { uint32_t count, marker, base; };       // marker: position of the first marker in the tile or 0xffffffff

// b[0] = byte before my 16 (0 at the segment start), b[1..16] = my bytes (0 past the end); keep: bit i = byte i is data
// (not the 00 of an FF00 pair); marker_at: first FF followed by a non-zero byte (or FF as the very last byte)
__device__ __forceinline__ void ju_classify(const uint8_t* __restrict__ in, const LongSeg& sg, uint32_t p0, uint8_t (&b)[17],
                                            uint32_t& keep, uint32_t& marker_at)
{
    // (tiles start at in_start rounded down to 16: the window of a thread is one aligned vector; bytes before
    // in_start are not part of the segment)
    b[0] = (p0 > sg.in_start && p0 - 1 < sg.in_end) ? in[p0 - 1] : 0;
    if (p0 + 16 <= sg.in_end && ((uintptr_t)(in + p0) & 15) == 0) {
        const uint4 v = __ldg((const uint4*)(in + p0));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 16; ++i) b[i + 1] = (uint8_t)(w[i >> 2] >> ((i & 3) * 8));
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) b[i + 1] = (p0 + i < sg.in_end) ? in[p0 + i] : 0;
    }
    keep = 0; marker_at = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t p = p0 + i;
        if (p < sg.in_end && p >= sg.in_start) {
            const bool stuffed = p > sg.in_start && b[i] == 0xFF && b[i + 1] == 0x00;
            if (!stuffed) keep |= 1u << i;
            if (b[i + 1] == 0xFF) {
                const uint32_t nx = (p + 1 < sg.in_end) ? (i < 15 ? b[i + 2] : in[p + 1]) : 0xFFu;
                if (nx != 0x00 && marker_at == 0xffffffffu) marker_at = p;
            }
        }
    }
}
__device__ __forceinline__ uint32_t ju_find_seg(const LongSeg* __restrict__ segs, int nsegs, uint32_t tile)
{
    int lo = 0, hi = nsegs - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (segs[mid].tile_base <= tile) lo = mid; else hi = mid - 1; }
    return (uint32_t)lo;
}

__global__ void __launch_bounds__(256)
jpeg_unstuff_count_kernel(const JpegImage* __restrict__ images, const LongSeg* __restrict__ segs, int nsegs, JuTile* __restrict__ tiles)
{
    __shared__ uint32_t s_cnt[8], s_mk[8];
    const LongSeg sg = segs[ju_find_seg(segs, nsegs, blockIdx.x)];
    const uint8_t* in = images[sg.image].data;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t p0 = (sg.in_start & ~15u) + (blockIdx.x - sg.tile_base) * (uint32_t)JU_TILE + tid * 16;
    uint8_t b[17]; uint32_t keep, mk;
    ju_classify(in, sg, p0, b, keep, mk);
    uint32_t cnt = __popc(keep);
#pragma unroll
    for (int d = 16; d; d >>= 1) { cnt += __shfl_xor_sync(0xffffffffu, cnt, d); mk = min(mk, __shfl_xor_sync(0xffffffffu, mk, d)); }
    if (lane == 0) { s_cnt[warp] = cnt; s_mk[warp] = mk; }
    __syncthreads();
    if (tid == 0) {
        uint32_t c = 0, m = 0xffffffffu;
        for (int w = 0; w < 8; ++w) { c += s_cnt[w]; m = min(m, s_mk[w]); }
        tiles[blockIdx.x].count = c; tiles[blockIdx.x].marker = m;
    }
}

__global__ void __launch_bounds__(32)
jpeg_unstuff_scan_kernel(const LongSeg* __restrict__ segs, JuTile* __restrict__ tiles, uint32_t* __restrict__ seg_end)
{
    const LongSeg sg = segs[blockIdx.x];
    const int lane = threadIdx.x;
    JuTile* T = tiles + sg.tile_base;
    uint32_t run = 0, endp = 0xffffffffu;
    for (uint32_t t0 = 0; t0 < sg.ntiles && endp == 0xffffffffu; t0 += 32) {
        const uint32_t t = t0 + lane;
        const uint32_t c = t < sg.ntiles ? T[t].count : 0u, m = t < sg.ntiles ? T[t].marker : 0xffffffffu;
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += n; }
        if (t < sg.ntiles) T[t].base = run + incl - c;
        run += __shfl_sync(0xffffffffu, incl, 31);
        uint32_t mm = m;
#pragma unroll
        for (int d = 16; d; d >>= 1) mm = min(mm, __shfl_xor_sync(0xffffffffu, mm, d));
        endp = mm;           // tiles are in stream order: the first group with a marker has the first marker
    }
    if (lane == 0) seg_end[blockIdx.x] = endp;
}

__global__ void __launch_bounds__(256)
jpeg_unstuff_write_kernel(const JpegImage* __restrict__ images, const LongSeg* __restrict__ segs, int nsegs, const JuTile* __restrict__ tiles,
                          const uint32_t* __restrict__ seg_end, uint8_t* __restrict__ clean, uint32_t* __restrict__ clean_len)
{
    __shared__ uint32_t warp_sums[8];
    __shared__ __align__(16) uint8_t s_out[JU_TILE + 16];
    const uint32_t si = ju_find_seg(segs, nsegs, blockIdx.x);
    const LongSeg sg = segs[si];
    const uint32_t endp = seg_end[si];
    const uint32_t t0 = (sg.in_start & ~15u) + (blockIdx.x - sg.tile_base) * (uint32_t)JU_TILE;
    if (t0 > endp) return;                                     // behind the end of the data (CTA-uniform)
    const uint8_t* in = images[sg.image].data;
    uint8_t* out = clean + sg.clean_off;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t p0 = t0 + tid * 16;
    uint8_t b[17]; uint32_t keep, mk;
    ju_classify(in, sg, p0, b, keep, mk);
    // bytes at or after the marker are not data
#pragma unroll
    for (int i = 0; i < 16; ++i) if (p0 + i >= endp) keep &= ~(1u << i);
    const uint32_t cnt = __popc(keep);
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += n; }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const uint32_t c = warp_sums[w]; woff += w < warp ? c : 0; total += c; }
    const uint32_t base = tiles[blockIdx.x].base;
    // the tile's output is put together in shared memory, at the byte alignment it has in memory (the stream is
    // 16-byte aligned), and leaves as whole words; the bytes at its two ends share words with the neighbouring tiles
    const uint32_t mis = base & 3u;
    uint32_t o = mis + woff + inc - cnt;
#pragma unroll
    for (int i = 0; i < 16; ++i) if (keep & (1u << i)) s_out[o++] = b[i + 1];
    __syncthreads();
    {
        const uint32_t lo = mis, hi = mis + total;              // bytes [lo, hi) of s_out <-> out[base - mis + ...]
        uint8_t* const g = out + (base - mis);
        const uint32_t wlo = (lo + 3) & ~3u, whi = hi & ~3u;
        if (wlo <= whi) {
            for (uint32_t w = wlo + tid * 4; w < whi; w += 256 * 4) *(uint32_t*)(g + w) = *(const uint32_t*)(s_out + w);
            if ((uint32_t)tid < wlo - lo) g[lo + tid] = s_out[lo + tid];
            if ((uint32_t)tid < hi - whi) g[whi + tid] = s_out[whi + tid];
        } else if ((uint32_t)tid < hi - lo) g[lo + tid] = s_out[lo + tid];
    }
    // the tile that holds the end of the data (the marker, or the last byte of the segment): length and padding --
    // reads past the end of the data return FF (jpegload.d:683-696)
    const bool last = endp != 0xffffffffu ? (endp - t0 < (uint32_t)JU_TILE) : (blockIdx.x - sg.tile_base == sg.ntiles - 1);
    if (last) {
        for (uint32_t i = tid; i < JS_PAD_BYTES; i += 256) out[base + total + i] = 0xFF;
        if (tid == 0) clean_len[si] = base + total;
    }
}

// ---- bit reader over the unstuffed stream (big-endian bit order) -----------------------------------------------
struct JsBits {
    // The word a refill needs was requested at the refill before it: the load's latency (the lanes of a warp read 32
    // different lines) is off the decode chain.
    const uint32_t* __restrict__ w; uint32_t next, ahead; uint64_t buf; int cnt;      // cnt valid bits, MSB-aligned
    __device__ __forceinline__ static uint32_t be(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
    __device__ __forceinline__ void init(const uint32_t* words, uint32_t bitpos)
    {
        w = words; next = bitpos >> 5;
        buf = ((uint64_t)be(__ldg(w + next)) << 32) | be(__ldg(w + next + 1));
        next += 2;
        ahead = __ldg(w + next);
        const int sh = bitpos & 31;
        buf <<= sh; cnt = 64 - sh;
    }
    __device__ __forceinline__ void refill() { if (cnt <= 32) { buf |= (uint64_t)be(ahead) << (32 - cnt); ++next; cnt += 32; ahead = __ldg(w + next); } }
    __device__ __forceinline__ uint32_t pos() const { return next * 32u - (uint32_t)cnt; }
    __device__ __forceinline__ uint32_t top32() const { return (uint32_t)(buf >> 32); }
    __device__ __forceinline__ void drop(int n) { buf <<= n; cnt -= n; }
};

struct JsImageCtx {            // per-CTA copy of what the decoders need from JpegImage
    int bpm; int comp_of[10]; int tab[6];     // tab[comp * 2 + (is AC)] = index into the global HuffTable array
    uint32_t comp_packed;                     // comp_of, two bits per block of the MCU
};

// Canonical decode of the 10..16-bit codes of a table with more long-code prefixes than sub-tables; result in the
// format of a table entry (len 0 = not a code word).
__device__ __noinline__ uint32_t js_slow_symbol(const HuffTable* __restrict__ h, uint32_t top)
{
    const uint32_t top16 = top >> 16;
    int sym = -1, len = 0;
    for (int l = HT_L1_BITS + 1; l <= 16; ++l) {
        const int code = (int)(top16 >> (16 - l));
        if (code <= h->maxcode[l] && code >= h->mincode[l]) { len = l; sym = h->val[h->valptr[l] + code - h->mincode[l]]; break; }
    }
    if (sym >= 0 && h->is_dc && sym > 15) sym = -1;
    if (sym < 0) return (uint32_t)(h->is_dc ? 1 : HT_KINC_EOB) << 9;
    const int size = sym & 15, run = sym >> 4;
    return (uint32_t)(len | (size << 5) | ((h->is_dc ? 1 : (size ? run + 1 : (run == 15 ? 16 : HT_KINC_EOB))) << 9));
}

// One Huffman symbol at the head of the bit buffer (>= 32 valid bits) through table t = comp * 2 + (is AC): the table
// entry (len | size << 5 | kinc << 9, len 0 = the bits are no code word). The second-level lookup is predicated,
// not a branch: the lanes of a warp are at different symbols.
__device__ __forceinline__ uint32_t js_symbol(uint32_t tabs, int t, const JsImageCtx& cx, const HuffTable* __restrict__ tables, uint32_t top)
{
    const uint32_t tb = tabs + (uint32_t)t * JS_TBL_BYTES;
    const uint32_t e1 = js_lds16(tb + ((top >> (32 - HT_L1_BITS)) << 1));
    uint32_t e = e1;
    if (e1 & HT_LONG) e = js_lds16(tb + JS_L1_BYTES + (((min(e1 & 7u, (uint32_t)HT_SUBS - 1u) << HT_L2_BITS) | ((top >> (32 - HT_L1_BITS - HT_L2_BITS)) & ((1u << HT_L2_BITS) - 1u))) << 1));
    if ((e1 & (HT_LONG | 7u)) == (HT_LONG | 7u)) e = js_slow_symbol(tables + cx.tab[t], top);       // rare
    return e;
}

// Same symbol through the two-level tables in global memory (used where threads of one CTA serve different images).
__device__ __forceinline__ uint32_t js_symbol_global(const HuffTable* __restrict__ h, uint32_t top)
{
    uint32_t e = h->l1[top >> (32 - HT_L1_BITS)];
    if (e & HT_LONG) {
        if ((e & 7u) == 7u) return js_slow_symbol(h, top);
        e = h->l2[e & 7][(top >> (32 - HT_L1_BITS - HT_L2_BITS)) & ((1 << HT_L2_BITS) - 1)];
    }
    return e;
}

// the value of `size` extra bits at the head of x (left-aligned), JPGD_HUFF_EXTEND (jpegload.d:816-822) without a branch:
// a leading 1 bit means the value is positive
__device__ __forceinline__ int js_extend(uint32_t x, uint32_t size)
{
    const uint32_t v = __funnelshift_l(x, 0u, size);                   // the top `size` bits of x, 0 for size 0
    return (int)v - (int)(((1u << size) - 1u) & ~(uint32_t)((int)x >> 31));
}

__device__ __forceinline__ void js_load_tables(FastTables& ft, JsImageCtx& cx, const JpegImage& im, const HuffTable* __restrict__ tables, int tid, int nthreads)
{
    if (tid == 0) {
        cx.bpm = im.blocks_per_mcu;
        uint32_t pk = 0;
        for (int b = 0; b < 10; ++b) { cx.comp_of[b] = b < im.blocks_per_mcu ? im.mcu_org[b] : 0; pk |= (uint32_t)(cx.comp_of[b] & 3) << (2 * b); }
        cx.comp_packed = pk;
        for (int c = 0; c < 3; ++c) { cx.tab[c * 2] = im.dc_tab[c < im.comps ? c : 0]; cx.tab[c * 2 + 1] = im.ac_tab[c < im.comps ? c : 0]; }
    }
    // 16-byte copies of l1 + l2 (2.5 KB) of every table in use; unused slots are never indexed
    const int comps = im.comps;
    constexpr int V = (int)(JS_TBL_BYTES / 16);
    for (int i = tid; i < 6 * V; i += nthreads) {
        const int t = i / V, r = i - t * V;
        const int c = t >> 1;
        if (c >= comps) continue;
        const HuffTable* h = tables + ((t & 1) ? im.ac_tab[c] : im.dc_tab[c]);
        ((uint4*)(ft.b + (size_t)t * JS_TBL_BYTES))[r] = __ldg((const uint4*)h->l1 + r);
    }
}

// Decodes the symbols that START in [st.bitpos, limit) without storing anything: exit state, blocks completed, and
// the sum of the DC differences per component. One straight-line body per symbol whatever its kind (the lanes of a
// warp are at different symbols): the table entry says how far the zig-zag index moves, a DC difference is kept until
// its block ends (or the chunk does) and added to its component's sum there.
__device__ __forceinline__ void js_scan_chunk(const uint32_t* __restrict__ words, ChunkState& st, uint32_t limit,
                                              uint32_t tabs, const JsImageCtx& cx, const HuffTable* __restrict__ tables,
                                              uint32_t& nblocks, int dcs[3])
{
    int bi = st.bik & 0xffff, k = st.bik >> 16;
    const int bpm = cx.bpm;
    const uint32_t cpk = cx.comp_packed;
    int comp = (int)((cpk >> (2 * bi)) & 3u);
    uint32_t nb = 0;
    int d0 = 0, d1 = 0, d2 = 0, cur_dc = 0;
    JsBits br; br.init(words, st.bitpos);
    uint32_t pos = st.bitpos;
    while (pos < limit) {
        br.refill();
        const bool is_dc = k == 0;
        const uint32_t top = br.top32();
        const uint32_t e = js_symbol(tabs, comp * 2 + (is_dc ? 0 : 1), cx, tables, top);
        uint32_t len = e & 31u;
        const uint32_t size = (e >> 5) & 15u;
        len += len == 0 ? 1u : 0u;                               // no code word: one bit, and the decoders keep moving
        const int diff = js_extend(top << len, size);
        cur_dc = is_dc ? diff : cur_dc;
        k += (int)((e >> 9) & 63u);
        br.drop((int)(len + size));
        pos += len + size;
        const bool end = k >= 64;
        d0 += (end && comp == 0) ? cur_dc : 0; d1 += (end && comp == 1) ? cur_dc : 0; d2 += (end && comp == 2) ? cur_dc : 0;
        cur_dc = end ? 0 : cur_dc;
        nb += end ? 1u : 0u;
        k = end ? 0 : k;
        const int nbi = bi + 1 == bpm ? 0 : bi + 1;
        bi = end ? nbi : bi;
        comp = (int)((cpk >> (2 * bi)) & 3u);
    }
    // the DC difference of a block in progress belongs to this chunk if its DC symbol started here
    d0 += comp == 0 ? cur_dc : 0; d1 += comp == 1 ? cur_dc : 0; d2 += comp == 2 ? cur_dc : 0;
    st.bitpos = pos; st.bik = (uint32_t)bi | ((uint32_t)k << 16);
    nblocks = nb; dcs[0] = d0; dcs[1] = d1; dcs[2] = d2;
}

// ---- 2. sync: CTA-local relaxation ------------------------------------------------------------------------------
// After the first pass only the chunks whose entry state is stale decode again. They are compacted every round so that
// the stale chunks of the whole CTA fill as few warps as possible (a warp costs a full chunk decode however few of
// its lanes work).
__global__ void __launch_bounds__(JS_CTA)
jpeg_sync_kernel(const JpegImage* __restrict__ images, const LongSeg* __restrict__ segs, int nsegs,
                 const uint8_t* __restrict__ clean, const uint32_t* __restrict__ clean_len,
                 const HuffTable* __restrict__ tables, ChunkRec* __restrict__ recs, ChunkState* __restrict__ entry_used)
{
    __shared__ FastTables ft;
    __shared__ JsImageCtx cx;
    __shared__ uint2 s_entry[JS_CTA], s_exit[JS_CTA], s_todo_entry[JS_CTA];
    __shared__ uint4 s_res[JS_CTA];                 // blocks completed, DC sums
    __shared__ uint16_t s_todo[JS_CTA];
    __shared__ uint32_t s_wcount[JS_CTA / 32];
    const uint32_t si = js_find_seg(segs, nsegs, blockIdx.x);
    const LongSeg sg = segs[si];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    js_load_tables(ft, cx, images[sg.image], tables, tid, JS_CTA);
    const uint32_t tabs = (uint32_t)__cvta_generic_to_shared(ft.b);
    const uint32_t local_cta = blockIdx.x - sg.cta_base;
    // chunk of slot s: the first JS_WARM slots re-decode the tail of the previous CTA's range
    const long long lc_first = (long long)local_cta * JS_OWN - JS_WARM;
    const uint32_t total_bits = clean_len[si] * 8;
    const uint32_t* words = (const uint32_t*)(clean + sg.clean_off);
    // chunks are laid out for the stuffed length; those wholly behind the unstuffed data hold nothing
    auto slot_active = [&](int s) {
        const long long l = lc_first + s;
        return l >= 0 && l < (long long)sg.nchunks && (l == 0 || (uint32_t)l * (uint32_t)JS_CHUNK_BITS < total_bits);
    };
    __syncthreads();

    {   // first pass. A chunk's entry state is not known, but a decoder that starts one chunk early on the guess "a
        // block starts at the first bit" has usually fallen into step with the true decode by the time it reaches the
        // chunk (Huffman codes self-synchronise): the run-up costs as much as the relaxation round it replaces and
        // leaves far fewer chunks with a wrong exit state for the rounds after it.
        const bool active = slot_active(tid);
        const uint32_t lc = active ? (uint32_t)(lc_first + tid) : 0;
        ChunkState st; st.bitpos = lc * JS_CHUNK_BITS; st.bik = 0;
        uint32_t nb = 0; int dcs[3] = {0, 0, 0};
        if (active && lc > 0) {
            st.bitpos = lc * JS_CHUNK_BITS - JS_RUNUP_BITS;
            js_scan_chunk(words, st, min(lc * (uint32_t)JS_CHUNK_BITS, total_bits), tabs, cx, tables, nb, dcs);
        }
        s_entry[tid] = make_uint2(st.bitpos, st.bik);
        nb = 0; dcs[0] = dcs[1] = dcs[2] = 0;
        if (active) js_scan_chunk(words, st, min((lc + 1) * (uint32_t)JS_CHUNK_BITS, total_bits), tabs, cx, tables, nb, dcs);
        else { st.bitpos = total_bits; st.bik = 0; }
        s_exit[tid] = make_uint2(st.bitpos, st.bik);
        s_res[tid] = make_uint4(nb, (uint32_t)dcs[0], (uint32_t)dcs[1], (uint32_t)dcs[2]);
    }
    // a chunk takes its predecessor's exit state unless it sits in slot 0 (whose guess stands) or is chunk 0 of the
    // segment (whose "guess" is the true start state)
    const bool chained = tid > 0 && slot_active(tid) && lc_first + tid > 0;
    for (;;) {
        __syncthreads();
        bool stale = false;
        uint2 prev = make_uint2(0, 0);
        if (chained) { prev = s_exit[tid - 1]; const uint2 mine = s_entry[tid]; stale = prev.x != mine.x || prev.y != mine.y; }
        const uint32_t bal = __ballot_sync(0xffffffffu, stale);
        if (lane == 0) s_wcount[warp] = __popc(bal);
        __syncthreads();
        uint32_t off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < JS_CTA / 32; ++w) { const uint32_t c = s_wcount[w]; off += w < warp ? c : 0; total += c; }
        if (total == 0) break;
        if (stale) { const uint32_t p = off + __popc(bal & ((1u << lane) - 1)); s_todo[p] = (uint16_t)tid; s_todo_entry[p] = prev; }
        __syncthreads();
        if ((uint32_t)tid < total) {
            const int s = s_todo[tid];
            const uint2 e = s_todo_entry[tid];
            const uint32_t lc = (uint32_t)(lc_first + s);
            ChunkState st; st.bitpos = e.x; st.bik = e.y;
            uint32_t nb = 0; int dcs[3];
            js_scan_chunk(words, st, min((lc + 1) * (uint32_t)JS_CHUNK_BITS, total_bits), tabs, cx, tables, nb, dcs);
            s_entry[s] = e;
            s_exit[s] = make_uint2(st.bitpos, st.bik);
            s_res[s] = make_uint4(nb, (uint32_t)dcs[0], (uint32_t)dcs[1], (uint32_t)dcs[2]);
        }
    }
    const long long lc_s = lc_first + tid;
    if (tid >= JS_WARM && lc_s < (long long)sg.nchunks) {
        const uint2 x = s_exit[tid]; const uint4 r = s_res[tid];
        ChunkRec* dst = recs + sg.chunk_base + (uint32_t)lc_s;
        ((uint4*)dst)[0] = make_uint4(x.x, x.y, r.x, r.y);
        ((uint4*)dst)[1] = make_uint4(r.z, r.w, 0u, 0u);
    }
    if (tid == JS_WARM) { ChunkState e; const uint2 v = s_entry[tid]; e.bitpos = v.x; e.bik = v.y; entry_used[blockIdx.x] = e; }   // the state assumed at the first own chunk
}

// ---- 3. repair / check of the CTA boundaries --------------------------------------------------------------------
// mode 0: repair (walk forward from a wrong boundary); mode 1: count boundaries that still disagree.
__global__ void __launch_bounds__(64)
jpeg_repair_kernel(const JpegImage* __restrict__ images, const LongSeg* __restrict__ segs, int nsegs, uint32_t total_ctas,
                   const uint8_t* __restrict__ clean, const uint32_t* __restrict__ clean_len,
                   const HuffTable* __restrict__ tables, ChunkRec* recs, ChunkState* entry_used, int mode, uint32_t* unconverged)
{
    const uint32_t cta = blockIdx.x * blockDim.x + threadIdx.x;
    if (cta >= total_ctas) return;
    const uint32_t si = js_find_seg(segs, nsegs, cta);
    const LongSeg sg = segs[si];
    const uint32_t local_cta = cta - sg.cta_base;
    if (local_cta == 0) return;                                  // starts from the true state
    const uint32_t lc0 = local_cta * JS_OWN;
    const uint32_t total_bits = clean_len[si] * 8;
    if (lc0 >= sg.nchunks || lc0 * (uint32_t)JS_CHUNK_BITS >= total_bits) return;     // nothing of its own to decode
    const ChunkRec* pr = recs + sg.chunk_base + lc0 - 1;
    const uint2 truth = __ldcg((const uint2*)pr);
    ChunkState used = entry_used[cta];
    if (truth.x == used.bitpos && truth.y == used.bik) return;
    if (mode == 1) { atomicAdd(unconverged, 1u); return; }
    // walk: taken by very few threads, through the global tables (threads of one CTA here serve different images)
    const JpegImage& im = images[sg.image];
    const uint32_t* words = (const uint32_t*)(clean + sg.clean_off);
    ChunkState st; st.bitpos = truth.x; st.bik = truth.y;
    entry_used[cta] = st;
    const uint32_t lc_end = min(lc0 + (uint32_t)JS_OWN, sg.nchunks);
    for (uint32_t lc = lc0; lc < lc_end; ++lc) {
        if (lc * (uint32_t)JS_CHUNK_BITS >= total_bits) break;
        const uint32_t limit = min((lc + 1) * (uint32_t)JS_CHUNK_BITS, total_bits);
        // plain decode through the global tables (same symbol semantics as js_scan_chunk)
        int bi = st.bik & 0xffff, k = st.bik >> 16;
        uint32_t nb = 0; int d[3] = {0, 0, 0};
        JsBits br; br.init(words, st.bitpos);
        uint32_t pos = st.bitpos;
        while (pos < limit) {
            br.refill();
            const int comp = im.mcu_org[bi];
            const bool is_dc = k == 0;
            const uint32_t top = br.top32();
            const uint32_t e = js_symbol_global(tables + (is_dc ? im.dc_tab[comp] : im.ac_tab[comp]), top);
            uint32_t len = e & 31u; const uint32_t size = (e >> 5) & 15u;
            if (len == 0) len = 1;
            if (is_dc) d[comp] += js_extend(top << len, size);
            k += (int)((e >> 9) & 63u);
            br.drop((int)(len + size)); pos += len + size;
            if (k >= 64) { k = 0; ++nb; bi = bi + 1 == im.blocks_per_mcu ? 0 : bi + 1; }
        }
        st.bitpos = pos; st.bik = (uint32_t)bi | ((uint32_t)k << 16);
        ChunkRec* dst = recs + sg.chunk_base + lc;
        const uint2 old = __ldcg((const uint2*)dst);
        ((uint4*)dst)[0] = make_uint4(st.bitpos, st.bik, nb, (uint32_t)d[0]);
        ((uint4*)dst)[1] = make_uint4((uint32_t)d[1], (uint32_t)d[2], 0u, 0u);
        if (old.x == st.bitpos && old.y == st.bik) break;        // met the old chain: everything after is already right
    }
}

// ---- 4. scan over the chunks of each segment: block base and DC bases (exclusive) -------------------------------
__global__ void __launch_bounds__(256)
jpeg_scan_kernel(const JpegImage* __restrict__ images, const LongSeg* __restrict__ segs, const ChunkRec* __restrict__ recs,
                 ChunkBase* __restrict__ bases, int* status)
{
    __shared__ uint32_t ws[8][4];
    __shared__ uint32_t carry[4];
    const LongSeg sg = segs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 4) carry[tid] = 0;
    __syncthreads();
    for (uint32_t t0 = 0; t0 < sg.nchunks; t0 += 256) {
        const uint32_t i = t0 + tid;
        uint32_t v[4] = {0, 0, 0, 0}, inc[4];
        if (i < sg.nchunks) {
            const ChunkRec* r = recs + sg.chunk_base + i;
            const uint4 a = ((const uint4*)r)[0]; const uint4 b = ((const uint4*)r)[1];
            v[0] = a.z; v[1] = a.w; v[2] = b.x; v[3] = b.y;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) inc[q] = v[q];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { const uint32_t n = __shfl_up_sync(0xffffffffu, inc[q], d); if (lane >= d) inc[q] += n; }
        }
        if (lane == 31) { ws[warp][0] = inc[0]; ws[warp][1] = inc[1]; ws[warp][2] = inc[2]; ws[warp][3] = inc[3]; }
        __syncthreads();
        uint32_t ex[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t off = carry[q];
            for (int w = 0; w < warp; ++w) off += ws[w][q];
            ex[q] = off + inc[q] - v[q];
            inc[q] += off;
        }
        if (i < sg.nchunks) *(uint4*)(bases + sg.chunk_base + i) = make_uint4(ex[0], ex[1], ex[2], ex[3]);
        __syncthreads();
        if (tid == 255) { carry[0] = inc[0]; carry[1] = inc[1]; carry[2] = inc[2]; carry[3] = inc[3]; }
        __syncthreads();
    }
    // the stream must hold every block of the segment
    if (tid == 0 && carry[0] < (uint32_t)sg.num_mcus * (uint32_t)images[sg.image].blocks_per_mcu) status[sg.image] = 0;
}

// ---- 5. write pass ------------------------------------------------------------------------------------------------
// Every lane assembles the block it is decoding in shared memory (64 int16 = 32 words; word w of lane l sits at
// [l][(w + l) & 31], so the lanes' scatters and the flush below are both free of bank conflicts). When a lane
// completes a block the whole warp writes it out -- one coalesced 128-byte row per block, zeros included -- and clears
// the lane's buffer: the coefficient array is written exactly once, in full sectors, with no zero-fill pass.
// (A lane writing its own block as eight 16-byte vectors was measured: 9.0 ms against 8.1; a CTA of 256 lanes taking
// 512 chunks from a shared counter, so that lanes with few symbols do not idle: 8.7 ms.)
constexpr int JW_CTA = 256;
__global__ void __launch_bounds__(JW_CTA)
jpeg_write_kernel(const JpegImage* __restrict__ images, const LongSeg* __restrict__ segs, int nsegs,
                  const uint8_t* __restrict__ clean, const uint32_t* __restrict__ clean_len,
                  const HuffTable* __restrict__ tables, const ChunkRec* __restrict__ recs, const ChunkBase* __restrict__ bases,
                  int* status)
{
    __shared__ FastTables ft;
    __shared__ JsImageCtx cx;
    __shared__ int16_t s_quant[3][64];
    __shared__ uint8_t s_zag[64];
    __shared__ __align__(16) uint32_t s_blk[JW_CTA][32];
    // CTA -> (segment, JW_CTA consecutive chunks): chunk_base of every segment is a multiple of JW_CTA (host)
    const uint32_t si = js_find_seg_chunk(segs, nsegs, blockIdx.x * JW_CTA);
    const LongSeg sg = segs[si];
    const JpegImage& im = images[sg.image];
    const int tid = threadIdx.x, lane = tid & 31;
    js_load_tables(ft, cx, im, tables, tid, JW_CTA);
    const uint32_t tabs = (uint32_t)__cvta_generic_to_shared(ft.b);
    for (int i = tid; i < 192; i += JW_CTA) s_quant[i >> 6][i & 63] = (i >> 6) < im.comps ? im.quant[i >> 6][i & 63] : (int16_t)0;
    if (tid < 64) s_zag[tid] = c_zag[tid];
#pragma unroll
    for (int w = 0; w < 32; ++w) s_blk[tid][(w + lane) & 31] = 0;
    __syncthreads();
    const uint32_t lc = blockIdx.x * JW_CTA + tid - sg.chunk_base;
    const uint32_t total_bits = clean_len[si] * 8;
    bool alive = lc < sg.nchunks && (lc == 0 || lc * (uint32_t)JS_CHUNK_BITS < total_bits);      // else: behind the data
    const uint32_t limit = min((lc + 1) * (uint32_t)JS_CHUNK_BITS, total_bits);
    const uint32_t hard_limit = total_bits + (JS_PAD_BYTES - 16) * 8;            // never read past the padding
    ChunkState st; st.bitpos = 0; st.bik = 0;
    uint4 bs = make_uint4(0, 0, 0, 0);
    if (alive) {
        if (lc > 0) { const uint2 v = __ldg((const uint2*)(recs + sg.chunk_base + lc - 1)); st.bitpos = v.x; st.bik = v.y; }
        bs = __ldg((const uint4*)(bases + sg.chunk_base + lc));
    }
    const int bpm = cx.bpm;
    const uint32_t seg_blocks = (uint32_t)sg.num_mcus * (uint32_t)bpm;
    int bi = st.bik & 0xffff, k = st.bik >> 16;
    int comp = cx.comp_of[bi];
    uint32_t b = bs.x;                              // index of the block in progress / about to start
    int dc0 = (int)bs.y, dc1 = (int)bs.z, dc2 = (int)bs.w;
    int16_t* const coef_seg = im.coefs + (size_t)sg.first_mcu * bpm * 64;
    uint8_t* const zag_seg = im.blk_zag + (size_t)sg.first_mcu * bpm;
    const uint32_t* words = (const uint32_t*)(clean + sg.clean_off);
    JsBits br; br.init(words, alive ? st.bitpos : 0);
    uint32_t pos = st.bitpos;
    bool mine = k == 0;                             // a block in progress at the entry belongs to an earlier chunk
    int last_k = 0;
    bool ok = true;
    uint32_t* const myrow = s_blk[tid];
    const uint32_t* const wrow = s_blk[tid & ~31];  // first row of my warp
    alive = alive && b < seg_blocks && pos < limit;
    // symbols of my blocks may run past `limit`; a new block is only started before it
    while (__any_sync(0xffffffffu, alive)) {
        bool flush = false;
        if (alive) {
            br.refill();
            const bool is_dc = k == 0;
            const uint32_t top = br.top32();
            const uint32_t e = js_symbol(tabs, comp * 2 + (is_dc ? 0 : 1), cx, tables, top);
            bool bad = (e & 31u) == 0;
            const int len = (int)(e & 31u) + (bad ? 1 : 0), size = (int)((e >> 5) & 15u), kinc = (int)((e >> 9) & 63u);
            int val = js_extend(top << len, (uint32_t)size);
            if (is_dc) {
                mine = true;
                val += comp == 0 ? dc0 : comp == 1 ? dc1 : dc2;
                dc0 = comp == 0 ? val : dc0; dc1 = comp == 1 ? val : dc1; dc2 = comp == 2 ? val : dc2;
                last_k = 0;
            } else if (size) {
                k += kinc - 1;                       // the run of zeros before the coefficient
                if (k > 63) { bad = true; k = 63; }
            } else if (kinc == 16 && k + 16 > 64) bad = true;
            if (mine && bad) ok = false;
            if (mine && (is_dc || size)) {
                const int zz = s_zag[k];
                ((int16_t*)(myrow + (((zz >> 1) + lane) & 31)))[zz & 1] = (int16_t)(val * s_quant[comp][k]);
                last_k = k;
            }
            k = is_dc ? 1 : (size ? k + 1 : (kinc == 16 ? k + 16 : 64));
            br.drop(len + size);
            pos += (uint32_t)(len + size);
            flush = k >= 64 && mine;
        }
        // whole-warp flush of the blocks completed in this round
        uint32_t fb = __ballot_sync(0xffffffffu, flush);
        if (fb) {
            __syncwarp();
            while (fb) {
                const int l = __ffs(fb) - 1; fb &= fb - 1;
                const uint32_t blk = __shfl_sync(0xffffffffu, b, l);
                uint32_t* row = (uint32_t*)wrow + l * 32 + ((lane + l) & 31);
                ((uint32_t*)(coef_seg + (size_t)blk * 64))[lane] = *row;
                *row = 0;
            }
            __syncwarp();
        }
        if (alive && k >= 64) {
            if (mine) zag_seg[b] = (uint8_t)(last_k + 1);
            k = 0; ++b; bi = bi + 1 == bpm ? 0 : bi + 1; comp = cx.comp_of[bi];
            if (pos >= limit) alive = false;        // the next block starts in a later chunk
        }
        alive = alive && (pos < limit || (mine && k != 0)) && pos < hard_limit && b < seg_blocks;
    }
    if (!ok) status[sg.image] = 0;
}
