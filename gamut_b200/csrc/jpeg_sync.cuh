// jpeg_sync.cuh -- intra-image parallel Huffman decoding for long entropy segments (included by jpeg.cu).
//
// A baseline JPEG scan without restart markers is one serial bit stream (decode_next_row,
// jpegload.d:2405-2525). Huffman codes self-synchronise, so the stream is cut into fixed-size chunks and
// decoded by one thread per chunk:
//   1. jpeg_unstuff_kernel   removes FF00 byte stuffing and stops at the first marker, giving a plain bit
//                            stream (padded with 1-bits: the reference reads all-ones past a marker,
//                            jpegload.d:683-743);
//   2. jpeg_sync_kernel      pass 0: every thread decodes its chunk from a guessed state (block 0 of the MCU,
//                            DC next) and records the decoder state (bit position, block-in-MCU, zig-zag
//                            index) at the first symbol that starts in the next chunk. Passes 1..n: a thread
//                            whose predecessor's exit state changed re-decodes from that state; the host
//                            repeats until no exit state changes. Chunk 0 starts from the true state, so the
//                            fixed point is the serial decode (induction over chunks);
//   3. jpeg_scan_kernel      exclusive scan over chunks: first block index and DC running sums per chunk;
//   4. jpeg_sync_kernel<W>   final pass: decode again and write dequantised AC coefficients and DC differences;
//   5. jpeg_dcfix_kernel     DC prediction: per-chunk walk adding the scanned base, dequantise DC.
#pragma once

constexpr int JS_CHUNK_BYTES = 128;
constexpr int JS_CHUNK_BITS = JS_CHUNK_BYTES * 8;
constexpr uint32_t JS_LONG_MIN = 1024;           // shorter segments are decoded by one thread each (jpeg_huffman_kernel)

struct LongSeg {
    int image;
    uint32_t in_start, in_end;      // raw (stuffed) byte range in the file
    uint32_t chunk_base, nchunks;   // this segment's slice of the per-chunk arrays (nchunks from the stuffed length)
    int first_mcu, num_mcus;
    unsigned long long clean_off;   // offset of the unstuffed stream in the clean arena (16-byte aligned)
};

struct ChunkState { uint32_t bitpos; uint16_t bi; uint16_t k; };   // k: 0 = DC next, 1..63 = next AC index
__device__ __forceinline__ ChunkState js_load(const ChunkState* p) { const uint2 v = __ldcg((const uint2*)p); ChunkState s; s.bitpos = v.x; s.bi = (uint16_t)(v.y & 0xffff); s.k = (uint16_t)(v.y >> 16); return s; }
__device__ __forceinline__ void js_store(ChunkState* p, ChunkState s) { __stcg((uint2*)p, make_uint2(s.bitpos, (uint32_t)s.bi | ((uint32_t)s.k << 16))); }
__device__ __forceinline__ uint32_t js_find_seg(const LongSeg* __restrict__ segs, int nsegs, uint32_t c)
{
    int lo = 0, hi = nsegs - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (segs[mid].chunk_base <= c) lo = mid; else hi = mid - 1; }
    return (uint32_t)lo;
}

// ---- 1. unstuff: one CTA per segment ---------------------------------------------------------------
__global__ void __launch_bounds__(256)
jpeg_unstuff_kernel(const JpegImage* __restrict__ images, const LongSeg* __restrict__ segs, uint8_t* __restrict__ clean,
                    uint32_t* __restrict__ clean_len)
{
    __shared__ uint32_t warp_sums[8];
    __shared__ uint32_t s_base, s_end;
    const LongSeg sg = segs[blockIdx.x];
    const uint8_t* in = images[sg.image].data;
    uint8_t* out = clean + sg.clean_off;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_base = 0; s_end = 0xffffffffu; }
    __syncthreads();
    uint32_t base = 0;
    for (uint32_t t0 = sg.in_start; t0 < sg.in_end; t0 += 256 * 16) {
        const uint32_t p0 = t0 + tid * 16;
        uint8_t b[17];
        // b[0] = byte before my 16 (0 at the segment start), b[1..16] = my bytes (0 past the end)
        b[0] = (p0 > sg.in_start && p0 - 1 < sg.in_end) ? in[p0 - 1] : 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) b[i + 1] = (p0 + i < sg.in_end) ? in[p0 + i] : 0;
        // a byte is dropped when it is the 00 of an FF00 pair; a marker is FF followed by a non-zero byte
        // (or FF as the very last byte of the data)
        uint32_t keep = 0, marker_at = 0xffffffffu;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t p = p0 + i;
            if (p < sg.in_end) {
                const bool stuffed = b[i] == 0xFF && b[i + 1] == 0x00;
                if (!stuffed) keep |= 1u << i;
                if (b[i + 1] == 0xFF) {
                    const uint32_t nx = (p + 1 < sg.in_end) ? (i < 15 ? b[i + 2] : in[p + 1]) : 0xFFu;
                    if (nx != 0x00 && marker_at == 0xffffffffu) marker_at = p;
                }
            }
        }
        // earliest marker in this tile
        uint32_t mk = marker_at;
#pragma unroll
        for (int d = 16; d; d >>= 1) mk = min(mk, __shfl_xor_sync(0xffffffffu, mk, d));
        if (lane == 0 && mk != 0xffffffffu) atomicMin(&s_end, mk);
        __syncthreads();
        const uint32_t endp = s_end;
        // bytes at or after the marker are not data
#pragma unroll
        for (int i = 0; i < 16; ++i) if (p0 + i >= endp) keep &= ~(1u << i);
        const uint32_t cnt = __popc(keep);
        uint32_t inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += n; }
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < warp; ++w) woff += warp_sums[w];
        uint32_t o = base + woff + inc - cnt;
#pragma unroll
        for (int i = 0; i < 16; ++i) if (keep & (1u << i)) out[o++] = b[i + 1];
        uint32_t tile_total = 0;
        for (int w = 0; w < 8; ++w) tile_total += warp_sums[w];
        base += tile_total;
        __syncthreads();
        if (endp != 0xffffffffu) break;
    }
    // pad with 1-bits: reads past the end of the data return FF (jpegload.d:683-696)
    for (uint32_t i = tid; i < 64; i += 256) out[base + i] = 0xFF;
    if (tid == 0) clean_len[blockIdx.x] = base;
}

// ---- 2/4. chunk decode ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t js_peek32(const uint32_t* __restrict__ words, uint32_t bitpos)
{
    const uint32_t idx = bitpos >> 5, sh = bitpos & 31;
    const uint32_t w0 = __byte_perm(words[idx], 0, 0x0123), w1 = __byte_perm(words[idx + 1], 0, 0x0123);
    return __funnelshift_l(w1, w0, sh);
}

// Decodes the symbols that start in [st.bitpos, limit). Returns false on a stream error (only meaningful
// when the entry state is the true state). WRITE: store coefficients of blocks [blk0, blk_end).
template <bool WRITE>
__device__ __forceinline__ bool js_decode(const uint32_t* __restrict__ words, ChunkState& st, uint32_t limit,
                                          const JpegImage& im, const HuffTable* __restrict__ tables,
                                          uint32_t& nblocks, int16_t* coef_seg, uint32_t blk0, uint32_t blk_end, int dcsum[3])
{
    uint32_t bitpos = st.bitpos; int bi = st.bi, k = st.k;
    const int bpm = im.blocks_per_mcu;
    uint32_t nb = 0;
    bool ok = true;          // errors count only while the current block lies inside [blk0, blk_end)
#define JS_ERR() do { if (blk0 + nb < blk_end) ok = false; } while (0)
    while (bitpos < limit) {
        const int comp = im.mcu_org[bi];
        const HuffTable* h = tables + (k == 0 ? im.dc_tab[comp] : im.ac_tab[comp]);
        const uint32_t v = js_peek32(words, bitpos);
        const uint32_t top = v >> 16;
        uint32_t e = h->fast[top >> (16 - HUFF_FAST)];
        int len, sym;
        if (e) { len = (int)(e >> 8); sym = (int)(e & 255); }
        else {
            len = 0; sym = -1;
            for (int l = HUFF_FAST + 1; l <= 16; ++l) {
                const int code = (int)(top >> (16 - l));
                if (code <= h->maxcode[l] && code >= h->mincode[l]) { len = l; sym = h->val[h->valptr[l] + code - h->mincode[l]]; break; }
            }
            if (sym < 0) { JS_ERR(); len = 1; sym = 0; }      // invalid code word: speculative decoders just move on
        }
        const int s = sym & 15;
        const uint32_t extra = s ? ((v << len) >> (32 - s)) : 0u;
        bitpos += (uint32_t)(len + s);
        if (k == 0) {
            if (sym > 15) JS_ERR();
            const int diff = huff_extend((int)extra, s);
            if (WRITE) { const uint32_t b = blk0 + nb; if (b < blk_end) coef_seg[(size_t)b * 64] = (int16_t)diff; dcsum[comp] += diff; }
            k = 1;
        } else {
            const int r = sym >> 4;
            if (s) {
                if (r) { if (k + r > 63) { JS_ERR(); k = 63; } else k += r; }
                if (WRITE) { const uint32_t b = blk0 + nb; if (b < blk_end) coef_seg[(size_t)b * 64 + c_zag[k]] = (int16_t)(huff_extend((int)extra, s) * im.quant[comp][k]); }
                k += 1;
            } else if (r == 15) { if (k + 16 > 64) { JS_ERR(); k = 64; } else k += 16; }
            else k = 64;
        }
        if (k >= 64) { k = 0; ++nb; bi = bi + 1 == bpm ? 0 : bi + 1; }
    }
#undef JS_ERR
    st.bitpos = bitpos; st.bi = (uint16_t)bi; st.k = (uint16_t)k;
    nblocks = nb;
    return ok;
}

// pass: 0 = speculative first pass, >0 = relaxation. exit/nblk are updated in place; `changed` counts updates.
__global__ void __launch_bounds__(128)
jpeg_sync_kernel(const JpegImage* __restrict__ images, const LongSeg* __restrict__ segs, int nsegs,
                 uint32_t total_chunks, const uint8_t* __restrict__ clean, const uint32_t* __restrict__ clean_len,
                 const HuffTable* __restrict__ tables, ChunkState* exitst, uint32_t* nblk,
                 const uint8_t* __restrict__ dirty_in, uint8_t* __restrict__ dirty_out, int pass, uint32_t* changed)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= total_chunks) return;
    const uint32_t si = js_find_seg(segs, nsegs, c);
    const LongSeg sg = segs[si];
    const uint32_t lc = c - sg.chunk_base;             // chunk index inside the segment
    const uint32_t total_bits = clean_len[si] * 8;
    // chunks are laid out for the stuffed length; those wholly behind the unstuffed data take no part
    if (lc > 0 && lc * (uint32_t)JS_CHUNK_BITS >= total_bits) {
        if (pass == 0) { ChunkState z; z.bitpos = total_bits; z.bi = 0; z.k = 0; js_store(exitst + c, z); nblk[c] = 0; }
        dirty_out[c] = 0;
        return;
    }
    if (pass > 0) {
        dirty_out[c] = 0;
        if (lc == 0 || !dirty_in[c - 1]) return;       // my entry state did not change
    }
    const JpegImage& im = images[sg.image];
    const uint32_t limit = min((lc + 1) * (uint32_t)JS_CHUNK_BITS, total_bits);
    ChunkState st;
    if (pass == 0 || lc == 0) { st.bitpos = lc * JS_CHUNK_BITS; st.bi = 0; st.k = 0; }
    if (pass > 0 && lc > 0) st = js_load(exitst + c - 1);
    uint32_t nb = 0; int dummy[3];
    js_decode<false>((const uint32_t*)(clean + sg.clean_off), st, limit, im, tables, nb, nullptr, 0, 0, dummy);
    if (pass == 0) { js_store(exitst + c, st); nblk[c] = nb; dirty_out[c] = 1; }
    else {
        const ChunkState old = js_load(exitst + c);
        const bool ch = old.bitpos != st.bitpos || old.bi != st.bi || old.k != st.k;
        js_store(exitst + c, st); nblk[c] = nb;
        if (ch) { dirty_out[c] = 1; atomicAdd(changed, 1u); }
    }
}

// ---- 3. scan over the chunks of each segment: block base (exclusive) -----------------------------------
__global__ void __launch_bounds__(256)
jpeg_scan_kernel(const LongSeg* __restrict__ segs, const uint32_t* __restrict__ vals, uint32_t* __restrict__ excl)
{
    __shared__ uint32_t ws[8];
    __shared__ uint32_t carry;
    const LongSeg sg = segs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (uint32_t t0 = 0; t0 < sg.nchunks; t0 += 256) {
        const uint32_t i = t0 + tid;
        const uint32_t v = i < sg.nchunks ? vals[sg.chunk_base + i] : 0;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += n; }
        if (lane == 31) ws[warp] = inc;
        __syncthreads();
        uint32_t off = carry;
        for (int w = 0; w < warp; ++w) off += ws[w];
        if (i < sg.nchunks) excl[sg.chunk_base + i] = off + inc - v;
        __syncthreads();
        if (tid == 255) carry = off + inc;
        __syncthreads();
    }
}
// same for the three DC sums (int32, wrap-around like the reference's uint accumulation)
__global__ void __launch_bounds__(256)
jpeg_scan3_kernel(const LongSeg* __restrict__ segs, const int* __restrict__ vals /* [chunk][3] */, int* __restrict__ excl)
{
    __shared__ int ws[8][3];
    __shared__ int carry[3];
    const LongSeg sg = segs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 3) carry[tid] = 0;
    __syncthreads();
    for (uint32_t t0 = 0; t0 < sg.nchunks; t0 += 256) {
        const uint32_t i = t0 + tid;
        int v[3], inc[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) { v[q] = i < sg.nchunks ? vals[(size_t)(sg.chunk_base + i) * 3 + q] : 0; inc[q] = v[q]; }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
            for (int q = 0; q < 3; ++q) { const int n = __shfl_up_sync(0xffffffffu, inc[q], d); if (lane >= d) inc[q] += n; }
        }
        if (lane == 31) { ws[warp][0] = inc[0]; ws[warp][1] = inc[1]; ws[warp][2] = inc[2]; }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            int off = carry[q];
            for (int w = 0; w < warp; ++w) off += ws[w][q];
            if (i < sg.nchunks) excl[(size_t)(sg.chunk_base + i) * 3 + q] = off + inc[q] - v[q];
            inc[q] += off;
        }
        __syncthreads();
        if (tid == 255) { carry[0] = inc[0]; carry[1] = inc[1]; carry[2] = inc[2]; }
        __syncthreads();
    }
}

// ---- 4. write pass ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
jpeg_write_kernel(const JpegImage* __restrict__ images, const LongSeg* __restrict__ segs, int nsegs,
                  uint32_t total_chunks, const uint8_t* __restrict__ clean, const uint32_t* __restrict__ clean_len,
                  const HuffTable* __restrict__ tables, const ChunkState* exitst, const uint32_t* __restrict__ blkbase,
                  int* __restrict__ dcsum, int* status)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= total_chunks) return;
    const uint32_t si = js_find_seg(segs, nsegs, c);
    const LongSeg sg = segs[si];
    const uint32_t lc = c - sg.chunk_base;
    const JpegImage& im = images[sg.image];
    const uint32_t total_bits = clean_len[si] * 8;
    const uint32_t limit = min((lc + 1) * (uint32_t)JS_CHUNK_BITS, total_bits);
    if (lc > 0 && lc * (uint32_t)JS_CHUNK_BITS >= total_bits) {      // behind the data
        dcsum[(size_t)c * 3 + 0] = 0; dcsum[(size_t)c * 3 + 1] = 0; dcsum[(size_t)c * 3 + 2] = 0;
        return;
    }
    ChunkState st;
    if (lc == 0) { st.bitpos = 0; st.bi = 0; st.k = 0; } else st = js_load(exitst + c - 1);
    const uint32_t seg_blocks = (uint32_t)sg.num_mcus * im.blocks_per_mcu;
    const uint32_t b0 = blkbase[c];
    int ds[3] = {0, 0, 0};
    uint32_t nb = 0;
    int16_t* coef_seg = im.coefs + (size_t)sg.first_mcu * im.blocks_per_mcu * 64;
    if (b0 < seg_blocks) {
        // symbols after the segment's last block (padding bits) are decoded but neither stored nor checked
        if (!js_decode<true>((const uint32_t*)(clean + sg.clean_off), st, limit, im, tables, nb, coef_seg, b0, seg_blocks, ds))
            status[sg.image] = 0;
    }
    dcsum[(size_t)c * 3 + 0] = ds[0]; dcsum[(size_t)c * 3 + 1] = ds[1]; dcsum[(size_t)c * 3 + 2] = ds[2];
    // the last chunk checks that the stream held all the blocks
    if ((lc + 1) * (uint32_t)JS_CHUNK_BITS >= total_bits && b0 + nb < seg_blocks) status[sg.image] = 0;
}

// ---- 5. DC prediction ------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
jpeg_dcfix_kernel(const JpegImage* __restrict__ images, const LongSeg* __restrict__ segs, int nsegs,
                  uint32_t total_chunks, const ChunkState* exitst, const uint32_t* __restrict__ blkbase,
                  const uint32_t* __restrict__ nblk, const int* __restrict__ dcbase)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= total_chunks) return;
    const uint32_t si = js_find_seg(segs, nsegs, c);
    const LongSeg sg = segs[si];
    const uint32_t lc = c - sg.chunk_base;
    const JpegImage& im = images[sg.image];
    const uint32_t seg_blocks = (uint32_t)sg.num_mcus * im.blocks_per_mcu;
    // blocks whose DC symbol was decoded by this chunk: from the first block that starts here ...
    const int entry_k = lc == 0 ? 0 : js_load(exitst + c - 1).k;
    const int exit_k = js_load(exitst + c).k;
    uint32_t b = blkbase[c] + (entry_k != 0 ? 1u : 0u);
    uint32_t e = blkbase[c] + nblk[c] + (exit_k != 0 ? 1u : 0u);      // ... to the one still open at the exit
    if (e > seg_blocks) e = seg_blocks;
    int dc[3] = {dcbase[(size_t)c * 3], dcbase[(size_t)c * 3 + 1], dcbase[(size_t)c * 3 + 2]};
    int16_t* coef_seg = im.coefs + (size_t)sg.first_mcu * im.blocks_per_mcu * 64;
    const int bpm = im.blocks_per_mcu;
    for (; b < e; ++b) {
        const int comp = im.mcu_org[b % bpm];
        int16_t* p = coef_seg + (size_t)b * 64;
        dc[comp] += (int)p[0];
        p[0] = (int16_t)(dc[comp] * im.quant[comp][0]);
    }
}
