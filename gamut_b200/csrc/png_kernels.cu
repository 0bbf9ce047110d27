// png_kernels.cu -- PNG device kernels: batched inflate, row unfilter (wavefront), finish.
#include "common.h"
#include "png_kernels.cuh"
#include "inflate_par.cuh"
#include <vector>
#include <algorithm>
#include <atomic>

namespace gb {

// ---------------------------------------------------------------------------------------------
// Batched inflate. Long streams go through the block-parallel pipeline of inflate_par.cuh; every stream that
// pipeline did not accept (short, odd, corrupt, output buffer too small) is decoded by one warp (inflate.cuh).
__global__ void __launch_bounds__(INF_WARPS_PER_CTA * 32)
inflate_batch_kernel(InflateJob* jobs, int njobs, const InfPar* par)
{
    __shared__ InflateSmem smem[INF_WARPS_PER_CTA];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int j = blockIdx.x * INF_WARPS_PER_CTA + warp;
    if (j >= njobs) return;
    if (par && par[j].eligible && par[j].ok && !par[j].fail) return;     // accepted by the parallel pipeline
    InflateJob job = jobs[j];
    inflate_stream(job, smem[warp], lane);
    if (lane == 0) { jobs[j].out_len = job.out_len; jobs[j].status = job.status; }
}

static std::atomic<int> g_inflate_mode{-1};      // -1: unset (env GB200_INFLATE: "serial" | "parallel"), 0 serial, 1 parallel
void set_inflate_mode(int m) { g_inflate_mode.store(m); }

bool launch_inflate(InflateJob* d_jobs, const InflateJob* h_jobs, int njobs, cudaStream_t st, InflateWork& W)
{
    if (njobs <= 0) return true;
    int mode = g_inflate_mode.load();
    if (mode < 0) {
        const char* e = getenv("GB200_INFLATE");
        mode = (e && !strcmp(e, "serial")) ? 0 : 1;
        g_inflate_mode.store(mode);
    }
    const int grid = (njobs + INF_WARPS_PER_CTA - 1) / INF_WARPS_PER_CTA;
    // ---- plan the parallel pipeline
    std::vector<InfPar> par((size_t)njobs);
    std::vector<uint32_t> tile_start((size_t)njobs + 1);
    size_t slots_total = 0, blocks_total = 0, bitmap_total = 0;
    uint64_t bits_total = 0;
    uint32_t tiles = 0;
    int neligible = 0;
    for (int j = 0; j < njobs; ++j) {
        const InflateJob& J = h_jobs[j];
        InfPar& P = par[j];
        memset(&P, 0, sizeof(P));
        tile_start[j] = tiles;
        const bool el = mode == 1 && J.in_len >= 2048 && J.in_len < (1u << 28) && J.out_cap >= 64 &&
                        J.out_cap < 0xfffffff0u && (((uintptr_t)J.out) & 3) == 0 && (((uintptr_t)J.in) & 15) == 0;
        if (!el) continue;
        P.eligible = 1;
        P.in_bits = J.in_len * 8;
        P.nslots = (P.in_bits + INFP_SLOT_BITS - 1) / INFP_SLOT_BITS;
        P.maxblocks = J.in_len / 2048 + 8;
        // offsets for now; turned into pointers below
        P.slots = (uint32_t*)(uintptr_t)slots_total;   slots_total += P.nslots;
        P.blocks = (InfBlock*)(uintptr_t)blocks_total; blocks_total += P.maxblocks;
        P.bitmap = (uint32_t*)(uintptr_t)bitmap_total; bitmap_total += ((size_t)(J.out_cap / 32) + 2 + 3) & ~(size_t)3;
        bits_total += P.in_bits;
        const uint32_t nwords = (J.in_len + 3) / 4;
        tiles += (nwords + INFP_TILE_WORDS - 1) / INFP_TILE_WORDS;
        ++neligible;
    }
    tile_start[njobs] = tiles;
    if (neligible == 0) {
        inflate_batch_kernel<<<grid, INF_WARPS_PER_CTA * 32, 0, st>>>(d_jobs, njobs, nullptr);
        count_launch();
        return true;
    }
    const uint32_t vq_cap = (uint32_t)std::min<uint64_t>(bits_total / 512 + 4096, 0x7fffffffull);
    const uint32_t vq2_cap = vq_cap / 4 + 1024;
    auto al256 = [](size_t n) { return (n + 255) & ~(size_t)255; };
    const size_t o_par = 0, o_tiles = al256(o_par + sizeof(InfPar) * njobs), o_ctr = al256(o_tiles + 4 * ((size_t)njobs + 1)),
                 o_slots = al256(o_ctr + 64), o_bitmap = al256(o_slots + 4 * slots_total), o_blocks = al256(o_bitmap + 4 * bitmap_total + 1024),
                 o_work = al256(o_blocks + sizeof(InfBlock) * blocks_total), o_vq = al256(o_work + 8 * blocks_total),
                 o_vq2 = al256(o_vq + 8 * (size_t)vq_cap), o_vq3 = al256(o_vq2 + 8 * (size_t)vq2_cap), total = al256(o_vq3 + 8 * (size_t)vq2_cap);
    if (!W.buf.alloc(total)) return false;
    uint8_t* base = W.buf.as<uint8_t>();
    for (int j = 0; j < njobs; ++j) {
        InfPar& P = par[j];
        if (!P.eligible) continue;
        P.slots = (uint32_t*)(base + o_slots) + (size_t)(uintptr_t)P.slots;
        P.blocks = (InfBlock*)(base + o_blocks) + (size_t)(uintptr_t)P.blocks;
        P.bitmap = (uint32_t*)(base + o_bitmap) + (size_t)(uintptr_t)P.bitmap;
    }
    // par and tile_start are uploaded from pageable memory: cudaMemcpyAsync returns once they are staged
    bool ok = true;
    ok &= cuda_ok(cudaMemcpyAsync(base + o_par, par.data(), sizeof(InfPar) * njobs, cudaMemcpyHostToDevice, st), "inflate par", __FILE__, __LINE__);
    ok &= cuda_ok(cudaMemcpyAsync(base + o_tiles, tile_start.data(), 4 * ((size_t)njobs + 1), cudaMemcpyHostToDevice, st), "inflate tiles", __FILE__, __LINE__);
    ok &= dev_fill_async(base + o_ctr, 0, 64, st);
    ok &= dev_fill_async(base + o_slots, 0xff, 4 * slots_total, st);
    ok &= dev_fill_async(base + o_bitmap, 0, 4 * bitmap_total, st);
    if (!ok) return false;
    InfPar* d_par = (InfPar*)(base + o_par);
    uint32_t* d_ctr = (uint32_t*)(base + o_ctr);
    uint2* d_work = (uint2*)(base + o_work);
    uint2* d_vq = (uint2*)(base + o_vq);
    uint2* d_vq2 = (uint2*)(base + o_vq2);
    uint2* d_vq3 = (uint2*)(base + o_vq3);
    const int persistent = sm_count() * 6;
    infp_find_kernel<<<tiles, 256, 0, st>>>(d_jobs, d_par, (const uint32_t*)(base + o_tiles), njobs, d_vq, vq_cap, d_ctr);
    // several passes with a growing symbol budget: most random headers over-subscribe their code within a few
    // symbols; whatever outlives a pass's budget is queued for the next pass and decoded again from its start
    // (keeps the lanes of a warp in step: without the budget a warp runs as long as its longest header)
    {
        const int budgets[4] = {12, 40, 120, 1 << 20};
        uint2* qin = d_vq; uint32_t qin_cap = vq_cap; uint32_t* cin = d_ctr + INFP_CTR_Q;
        for (int pass = 0; pass < 4; ++pass) {
            uint2* qout = (pass & 1) ? d_vq3 : d_vq2;
            uint32_t* cout = d_ctr + INFP_CTR_Q2 + pass;
            if (pass == 0)
                infp_verify_kernel<false><<<sm_count() * 8, 128, 0, st>>>(d_jobs, d_par, qin, qin_cap, cin, budgets[pass], qout, vq2_cap, cout);
            else
                infp_verify_kernel<true><<<sm_count() * 4, 128, 0, st>>>(d_jobs, d_par, qin, qin_cap, cin, budgets[pass],
                                                                         pass < 3 ? qout : nullptr, vq2_cap, pass < 3 ? cout : nullptr);
            qin = qout; qin_cap = vq2_cap; cin = cout;
        }
    }
    infp_compact_kernel<<<(njobs + 3) / 4, 128, 0, st>>>(d_par, njobs, d_work, d_ctr);
    infp_count_kernel<<<persistent, INFP_WARPS * 32, 0, st>>>(d_jobs, d_par, d_work, d_ctr);
    infp_walk_kernel<<<(njobs + INFP_WARPS - 1) / INFP_WARPS, INFP_WARPS * 32, 0, st>>>(d_jobs, d_par, njobs);
    infp_write_kernel<<<persistent, INFP_WARPS * 32, 0, st>>>(d_jobs, d_par, d_work, d_ctr);
    infp_resolve_kernel<<<njobs, LZ4C_NT, 0, st>>>(d_jobs, d_par, njobs);
    inflate_batch_kernel<<<grid, INF_WARPS_PER_CTA * 32, 0, st>>>(d_jobs, njobs, d_par);
    count_launch(11);
    return true;
}

// ---------------------------------------------------------------------------------------------
// Gather: concatenates the IDAT payloads of every image into one padded, 16-byte aligned stream per
// image (stbdec.d:1952-1989 does this with realloc + getn on the host).
struct Segment { const uint8_t* src; uint8_t* dst; uint32_t len; };

__global__ void __launch_bounds__(256)
gather_segments_kernel(const Segment* segs, int nsegs)
{
    // 16-byte destination vectors; the source is read as aligned words and realigned with a funnel shift (the up to
    // 3 bytes read past a segment are the chunk's CRC, inside the file)
    for (int s = blockIdx.x; s < nsegs; s += gridDim.x) {
        const Segment g = segs[s];
        uint32_t head = (16u - (uint32_t)((uintptr_t)g.dst & 15u)) & 15u;
        if (head > g.len) head = g.len;
        const uint32_t nvec = (g.len - head) >> 4;
        const uint32_t tail0 = head + (nvec << 4);
        if (threadIdx.x < head) g.dst[threadIdx.x] = g.src[threadIdx.x];
        if (tail0 + threadIdx.x < g.len) g.dst[tail0 + threadIdx.x] = g.src[tail0 + threadIdx.x];
        for (uint32_t v = threadIdx.x; v < nvec; v += 256) {
            const uint8_t* sp = g.src + head + ((size_t)v << 4);
            const uint32_t m = (uint32_t)((uintptr_t)sp & 3u), sh = m * 8;
            const uint32_t* wp = (const uint32_t*)(sp - m);
            const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3);
            const uint32_t w4 = m ? __ldg(wp + 4) : 0u;
            uint4 o;
            o.x = __funnelshift_r(w0, w1, sh); o.y = __funnelshift_r(w1, w2, sh);
            o.z = __funnelshift_r(w2, w3, sh); o.w = __funnelshift_r(w3, w4, sh);
            *(uint4*)(g.dst + head + ((size_t)v << 4)) = o;
        }
    }
}
void launch_gather(const void* d_segs, int nsegs, cudaStream_t st)
{
    if (nsegs <= 0) return;
    int grid = nsegs < 148 * 16 ? nsegs : 148 * 16;
    gather_segments_kernel<<<grid, 256, 0, st>>>((const Segment*)d_segs, nsegs);
    count_launch();
}

// ---------------------------------------------------------------------------------------------
// Row unfilter. PNG filters Sub/Avg/Paeth are serial along a row and Up/Avg/Paeth depend on the
// row above, so a job is processed as a skewed wavefront: lane l of a warp owns row (band*32 + l)
// and runs one pixel behind lane l-1; the pixel above comes from a warp shuffle, the pixel
// above-left is last step's shuffle result. With a zero row above row 0 and zero pixels left of
// column 0 the five filters reduce to the reference's first-row / first-pixel special cases
// (stbdec.d:1381-1388,1453-1465).
__device__ __forceinline__ int paeth_pred(int a, int b, int c)
{
    int p = a + b - c;
    int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

template <int BPP>
__device__ void unfilter_warp(const UnfilterJob& J, int* status, int lane)
{
    const uint32_t rb = J.row_bytes, H = J.height;
    const uint32_t npx = rb / BPP;
    for (uint32_t band = 0; band * 32 < H; ++band) {
        const uint32_t row = band * 32 + lane;
        const bool valid = row < H;
        const uint8_t* rr = J.raw + (size_t)(valid ? row : 0) * (rb + 1);
        int f = valid ? rr[0] : 0;
        if (f > 4) { status[J.image] = 0; f = 0; }       // "invalid filter" (stbdec.d:1438)
        rr += 1;
        uint8_t* orow = J.out + (size_t)(valid ? row : 0) * J.out_pitch;
        const uint8_t* prow = orow - J.out_pitch;
        uint64_t cur = 0, left = 0, upleft = 0;
        for (uint32_t t = 0; t < npx + 31; ++t) {
            const int x = (int)t - lane;
            uint64_t up;
            if (BPP <= 4) up = __shfl_up_sync(0xffffffffu, (uint32_t)cur, 1);
            else up = __shfl_up_sync(0xffffffffu, cur, 1);
            const bool active = valid && x >= 0 && x < (int)npx;
            if (lane == 0) {
                up = 0;
                if (active && row > 0) {
#pragma unroll
                    for (int k = 0; k < BPP; ++k) up |= (uint64_t)__ldcg(prow + (size_t)x * BPP + k) << (8 * k);
                }
            }
            if (active) {
                if (x == 0) { left = 0; upleft = 0; }
                uint64_t o = 0;
#pragma unroll
                for (int k = 0; k < BPP; ++k) {
                    int raw = rr[(size_t)x * BPP + k];
                    int a = (int)(left >> (8 * k)) & 255, b = (int)(up >> (8 * k)) & 255, c = (int)(upleft >> (8 * k)) & 255;
                    int pred = f == 0 ? 0 : f == 1 ? a : f == 2 ? b : f == 3 ? ((a + b) >> 1) : paeth_pred(a, b, c);
                    int v = (raw + pred) & 255;
                    o |= (uint64_t)v << (8 * k);
                    orow[(size_t)x * BPP + k] = (uint8_t)v;
                }
                cur = o;
                left = o;
            }
            upleft = up;
        }
        __threadfence_block();
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Row unfilter, fast path for 4-byte pixels (RGBA8 / LA16 / ...): one CTA of U4_NW warps per job.
//  * warp w owns bands w, w+U4_NW, ... (a band = 32 rows); inside a band the 32 lanes run the same
//    skewed wavefront as above, on whole pixels packed in one 32-bit register (SWAR byte arithmetic;
//    VABSDIFF4 is the only native byte-SIMD instruction on sm_100, the rest lowers to LOP3/IADD);
//  * the filtered rows are staged into shared memory with coalesced 4-byte cp.async (one 128-byte
//    transaction per row and chunk), three chunks ahead of use, in a 4-chunk ring per row; rows start at
//    arbitrary byte offsets (each row is prefixed by its filter byte), so the ring holds aligned words
//    and a pixel is extracted with one funnel shift;
//  * results go to a 2-chunk shared ring and are flushed with coalesced 128-byte stores per row;
//  * the row above a band's first row is the last row of the previous band, produced by another warp
//    of the same CTA: it is re-read from the (L2-resident) output through the same cp.async ring once
//    the producer has published the chunk in a shared progress counter.
// Shared-memory bank = (step - lane) mod 32 for every ring access: conflict-free by construction.
constexpr int U4_NW = 4;
constexpr int U4_CH = 16;           // wavefront steps (= skewed columns) per chunk
constexpr int U4_INW = 64;          // input ring: chunk in use + the next one (complete) + one in flight, of 16 columns
constexpr int U4_OUTW = 32;         // output ring: two chunks
// rows are padded by one word: every lane reads the same column in a step, so the row stride must be odd (bank = lane + column)
struct U4Smem { uint32_t in[33][U4_INW + 1]; uint32_t out[32][U4_OUTW + 1]; };     // 12.5 KB per warp => 16 warps per SM

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// Paeth predictor on four packed bytes (stbi__paeth, stbdec.d:1390-1401), 26 instructions: VABSDIFF4 is the only
// native byte-SIMD instruction on sm_100 and a bytewise compare costs 6, so the three-way comparison is reduced to two
// compares. With pa = |b-c|, pb = |a-c|, T = |a-b|, U = |pa-pb|: pc = |a+b-2c| equals pa+pb when a-c and b-c have the
// same sign (exactly when T == U; then pc >= pa, pb and the nearer of a, b wins) and U otherwise. So with q =
// min(pa, pb) and t = the value it belongs to (a on ties): result = (q <= pc) ? t : c, where pc may be replaced by
// 255 in the same-sign case. Checked against the reference formula for all 2^24 (a, b, c).
__device__ __forceinline__ uint32_t paeth4(uint32_t a, uint32_t b, uint32_t c)
{
    const uint32_t pa = __vabsdiffu4(b, c), pb = __vabsdiffu4(a, c);
    const uint32_t T = __vabsdiffu4(a, b), U = __vabsdiffu4(pa, pb);
    const uint32_t le = __vcmpleu4(pa, pb);
    const uint32_t q = (pa & le) | (pb & ~le);
    const uint32_t z = __vabsdiffu4(T, U);
    const uint32_t nz = ((z & 0x7f7f7f7fu) + 0x7f7f7f7fu) | z;          // bit 7 of a byte: that byte of z is non-zero
    uint32_t same;                                                      // replicate bit 7 over the byte (PRMT with the
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(same) : "r"(~nz & 0x80808080u), "r"(0u), "r"(0xba98u));   // sign-replicate selector bit; __byte_perm masks it off)
    const uint32_t ok = __vcmpleu4(q, U | same);
    const uint32_t t = (a & le) | (b & ~le);
    return (t & ok) | (c & ~ok);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gsrc) : "memory");
}

__device__ void unfilter4_cta(const UnfilterJob& J, int* status, U4Smem* S, volatile int* flushed, int warp, int lane)
{
    const uint32_t rb = J.row_bytes, H = J.height, npx = rb >> 2;
    const int nbands = (int)((H + 31) / 32);
    const int NSC = (int)((npx + 31 + U4_CH - 1) / U4_CH);   // step chunks per band (31 steps of skew)
    // flushed[w] = (bands finished by warp w) * npx + pixels of the current band's last row that are in memory
    const int pw = (warp + U4_NW - 1) % U4_NW;         // producer of the row above my bands
    const bool out16 = ((J.out_pitch & 15) == 0) && ((((uintptr_t)J.out) & 15) == 0);
    int kband = 0;
    for (int band = warp; band < nbands; band += U4_NW, ++kband) {
        const uint32_t r0 = (uint32_t)band * 32;
        const uint32_t row = r0 + lane;
        const bool valid = row < H;
        const uint8_t* rowaddr = J.raw + (size_t)(valid ? row : H - 1) * (rb + 1) + 1;
        int f = valid ? rowaddr[-1] : 0;
        if (f > 4) { status[J.image] = 0; f = 0; }            // "invalid filter" (stbdec.d:1438)
        const bool anyP = __any_sync(0xffffffffu, f == 4), anyA = __any_sync(0xffffffffu, f == 3);
        const int kprev = band > 0 ? (band - 1) / U4_NW : 0;
        const uint8_t* raw0 = J.raw + (size_t)r0 * (rb + 1) + 1;
        const uint8_t* bnd = band > 0 ? J.out + (size_t)(r0 - 1) * J.out_pitch : nullptr;
        const uint32_t nrows = min(32u, H - r0);

        if (!anyP && !anyA) {
            // ---- row-parallel mode: only None/Sub/Up rows => no serial dependency except Sub's prefix sum.
            // Each lane owns 4 consecutive pixels (16 bytes) of a 128-pixel column block; blocks are the outer
            // loop, rows the inner one, so the pixels above stay in registers. Loads are 16-byte vectors straight
            // from global memory (two per lane and row: rows start at arbitrary byte offsets), four rows in flight.
            const uint32_t maskU = __ballot_sync(0xffffffffu, f == 2), maskS = __ballot_sync(0xffffffffu, f == 1);
            const int NB = (int)((npx + 127) / 128);
            uint32_t carry = 0;                               // lane r: last output pixel of row r in the previous block
            uint32_t up0 = 0, up1 = 0, up2 = 0, up3 = 0;
            const int ngroups = (int)((nrows + 3) / 4);
            const int total = NB * ngroups;

            auto load_group = [&](int gi, uint4 (&A)[4], uint4 (&B)[4]) {
                // always load (addresses clamped into the band) so that no select has to wait for the data;
                // lanes / rows past the end read valid memory and their values are never stored
                gi = min(gi, total - 1);
                const int xb = gi / ngroups; const uint32_t y0 = (uint32_t)(gi - xb * ngroups) * 4;
                uint32_t px = (uint32_t)xb * 128 + lane * 4;
                if (px >= npx) px = 0;
                const uint8_t* rp = raw0 + (size_t)px * 4;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint8_t* q = rp + (size_t)min(y0 + u, nrows - 1) * (rb + 1);
                    const uint4* v = (const uint4*)(q - ((uintptr_t)q & 15));
                    A[u] = __ldg(v);
                    B[u] = __ldg(v + 1);
                }
            };
            auto compute_group = [&](int gi, const uint4 (&A)[4], const uint4 (&B)[4]) {
                const int xb = gi / ngroups; const uint32_t y0 = (uint32_t)(gi - xb * ngroups) * 4;
                const uint32_t px = (uint32_t)xb * 128 + lane * 4;
                if (y0 == 0) {                                // block start: the row above the band
                    up0 = up1 = up2 = up3 = 0;
                    if (bnd) {
                        const int need = kprev * (int)npx + (int)min(128u * (uint32_t)(xb + 1), npx);
                        while (flushed[pw] < need) __nanosleep(32);
                        const uint32_t* bp = (const uint32_t*)(bnd + (size_t)px * 4);
                        if (px + 3 < npx && out16) { const uint4 t = __ldcg((const uint4*)bp); up0 = t.x; up1 = t.y; up2 = t.z; up3 = t.w; }
                        else { if (px < npx) up0 = __ldcg(bp); if (px + 1 < npx) up1 = __ldcg(bp + 1); if (px + 2 < npx) up2 = __ldcg(bp + 2); if (px + 3 < npx) up3 = __ldcg(bp + 3); }
                    }
                }
                const uint8_t* rp = raw0 + (size_t)px * 4;
                uint8_t* o = J.out + (size_t)r0 * J.out_pitch + (size_t)px * 4;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t y = y0 + u;
                    if (y < nrows) {                          // warp-uniform
                        const uint8_t* q = rp + (size_t)y * (rb + 1);
                        const uint32_t m = (uint32_t)(uintptr_t)q & 15u, shb = (m & 3u) * 8u;
                        uint32_t w0, w1, w2, w3, w4;
                        switch (m >> 2) {                     // uniform: every lane of the row has the same misalignment
                        case 0: w0 = A[u].x; w1 = A[u].y; w2 = A[u].z; w3 = A[u].w; w4 = B[u].x; break;
                        case 1: w0 = A[u].y; w1 = A[u].z; w2 = A[u].w; w3 = B[u].x; w4 = B[u].y; break;
                        case 2: w0 = A[u].z; w1 = A[u].w; w2 = B[u].x; w3 = B[u].y; w4 = B[u].z; break;
                        default: w0 = A[u].w; w1 = B[u].x; w2 = B[u].y; w3 = B[u].z; w4 = B[u].w; break;
                        }
                        uint32_t v0 = __funnelshift_r(w0, w1, shb), v1 = __funnelshift_r(w1, w2, shb);
                        uint32_t v2 = __funnelshift_r(w2, w3, shb), v3 = __funnelshift_r(w3, w4, shb);
                        if ((maskS >> y) & 1u) {
                            v1 = __vadd4(v1, v0); v2 = __vadd4(v2, v1); v3 = __vadd4(v3, v2);
                            uint32_t inc = v3;
#pragma unroll
                            for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc = __vadd4(inc, n); }
                            uint32_t ex = __shfl_up_sync(0xffffffffu, inc, 1);
                            if (lane == 0) ex = 0;
                            ex = __vadd4(ex, __shfl_sync(0xffffffffu, carry, (int)y));
                            v0 = __vadd4(v0, ex); v1 = __vadd4(v1, ex); v2 = __vadd4(v2, ex); v3 = __vadd4(v3, ex);
                            // last pixel of the row inside this block = carry into the next block
                            const uint32_t lastpx = min(npx - 1, (uint32_t)xb * 128 + 127);
                            const uint32_t lsel = (lastpx & 3u) == 0 ? v0 : (lastpx & 3u) == 1 ? v1 : (lastpx & 3u) == 2 ? v2 : v3;
                            const uint32_t last = __shfl_sync(0xffffffffu, lsel, (int)((lastpx >> 2) & 31u));
                            if (lane == (int)y) carry = last;
                        } else {
                            const uint32_t mu = 0u - ((maskU >> y) & 1u);
                            v0 = __vadd4(v0, up0 & mu); v1 = __vadd4(v1, up1 & mu); v2 = __vadd4(v2, up2 & mu); v3 = __vadd4(v3, up3 & mu);
                            if (maskS) {                      // some other row of the band is Sub: keep its carry current
                                const uint32_t lastpx = min(npx - 1, (uint32_t)xb * 128 + 127);
                                const uint32_t lsel = (lastpx & 3u) == 0 ? v0 : (lastpx & 3u) == 1 ? v1 : (lastpx & 3u) == 2 ? v2 : v3;
                                const uint32_t last = __shfl_sync(0xffffffffu, lsel, (int)((lastpx >> 2) & 31u));
                                if (lane == (int)y) carry = last;
                            }
                        }
                        up0 = v0; up1 = v1; up2 = v2; up3 = v3;
                        uint32_t* op = (uint32_t*)(o + (size_t)y * J.out_pitch);
                        if (px + 3 < npx && out16) __stcs((uint4*)op, make_uint4(v0, v1, v2, v3));
                        else { if (px < npx) op[0] = v0; if (px + 1 < npx) op[1] = v1; if (px + 2 < npx) op[2] = v2; if (px + 3 < npx) op[3] = v3; }
                    }
                }
                if ((int)(y0 / 4) == ngroups - 1) {           // block end: publish it for the band below
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0) flushed[warp] = kband * (int)npx + (int)min(128u * (uint32_t)(xb + 1), npx);
                }
            };

            uint4 A0[4], B0[4];
            for (int gi = 0; gi < total; ++gi) {
                load_group(gi, A0, B0);
                compute_group(gi, A0, B0);
            }
            continue;
        }

        // ---- wavefront mode. Step t: lane l works on pixel x = t - l of its row. Both rings are indexed by the
        // *skewed* column t (row l's pixel x sits in column x + l), so every lane touches the same column in a step
        // (one shared address computation, bank = lane) and the rings only need to hold the chunk in use and the next.
        const uint32_t mis = (uint32_t)(uintptr_t)rowaddr & 3u, sh = mis * 8u;
        const uint32_t mS = f == 1 ? ~0u : 0u, mU = f == 2 ? ~0u : 0u, mA = f == 3 ? ~0u : 0u, mP = f == 4 ? ~0u : 0u;
        // staging / flushing: instruction k of a chunk handles rows 2k and 2k+1, 16 columns each (64-byte segments).
        // Row rr starts at raw0 + rr*(rb+1); rb is a multiple of 4, so its misalignment is (m0 + rr) & 3 and the
        // aligned word address of (row 2k+h, column col) is base[k&1] + k*kstep + 4*col with two lane constants.
        const uint32_t hrow = lane >> 4, jcol = lane & 15;
        const uint32_t m0 = (uint32_t)(uintptr_t)raw0 & 3u;
        const uint32_t e0 = (m0 + hrow) & 3u;
        const uint8_t* sbase0 = raw0 + (size_t)hrow * (rb + 1) - 4 * (size_t)hrow + 4 * (size_t)jcol - e0;
        const uint8_t* sbase1 = sbase0 + e0 - (e0 ^ 2u);
        const size_t kstep = 2 * (size_t)(rb + 1) - 8;
        uint32_t* const sdst = &S->in[hrow][jcol];
        uint8_t* const fbase = J.out + (size_t)(r0 + hrow) * J.out_pitch - 4 * (size_t)hrow + 4 * (size_t)jcol;
        const size_t fstep = 2 * (size_t)J.out_pitch - 8;
        const uint32_t* const fsrc = &S->out[hrow][jcol];
        auto stage = [&](int c) {
            if (c <= NSC) {                                   // one chunk past the last: the funnel reads column t + 1
                const int xb = c * U4_CH + (int)jcol - (int)hrow;             // pixel index for k = 0
                const uint32_t col = (uint32_t)(c * U4_CH) & (U4_INW - 1);
                const uint8_t* p0 = sbase0 + (size_t)c * (4 * U4_CH);
                const uint8_t* p1 = sbase1 + (size_t)c * (4 * U4_CH);
                const bool interior = c * U4_CH >= 32 && (uint32_t)((c + 1) * U4_CH) <= npx && nrows == 32;
                if (interior) {
#pragma unroll
                    for (uint32_t k = 0; k < 16; ++k)
                        cp_async4(sdst + 2 * k * (U4_INW + 1) + col, ((k & 1) ? p1 : p0) + k * kstep);
                } else {
#pragma unroll
                    for (uint32_t k = 0; k < 16; ++k) {
                        const int x = xb - 2 * (int)k;
                        if (2 * k + hrow < nrows && x >= 0 && (uint32_t)x <= npx)
                            cp_async4(sdst + 2 * k * (U4_INW + 1) + col, ((k & 1) ? p1 : p0) + k * kstep);
                    }
                }
                if (bnd) {
                    // the row above the band (skew 0): its producer must have it in memory up to this chunk
                    const uint32_t xe = min((uint32_t)(c + 1) * U4_CH, npx);
                    const int need = kprev * (int)npx + (int)xe;
                    while (flushed[pw] < need) __nanosleep(64);
                    const uint32_t x = (uint32_t)c * U4_CH + lane;
                    if (lane < U4_CH && x < npx) cp_async4(&S->in[32][x & (U4_INW - 1)], bnd + (size_t)x * 4);
                }
            }
            cp_async_commit();
        };
        auto flush = [&](int c) {
            const int xb = c * U4_CH + (int)jcol - (int)hrow;
            const uint32_t col = (uint32_t)(c * U4_CH) & (U4_OUTW - 1);
            uint8_t* p = fbase + (size_t)c * (4 * U4_CH);
            const bool interior = c * U4_CH >= 32 && (uint32_t)((c + 1) * U4_CH) <= npx && nrows == 32;
            if (interior) {
                uint32_t v[16];
#pragma unroll
                for (uint32_t k = 0; k < 16; ++k) v[k] = fsrc[2 * k * (U4_OUTW + 1) + col];
#pragma unroll
                for (uint32_t k = 0; k < 16; ++k) *(uint32_t*)(p + k * fstep) = v[k];
            } else {
#pragma unroll
                for (uint32_t k = 0; k < 16; ++k) {
                    const int x = xb - 2 * (int)k;
                    if (2 * k + hrow < nrows && x >= 0 && (uint32_t)x < npx) *(uint32_t*)(p + k * fstep) = fsrc[2 * k * (U4_OUTW + 1) + col];
                }
            }
            __threadfence_block();
            __syncwarp();
            // the band's last row (lane nrows-1) is in memory up to pixel 16(c+1) - (nrows-1)
            const int done = (c + 1) * U4_CH - (int)(nrows - 1);
            if (lane == 0) flushed[warp] = kband * (int)npx + (done < 0 ? 0 : done > (int)npx ? (int)npx : done);
        };

        stage(0);
        stage(1);
        uint32_t cur = 0, upleft = 0;
        const uint32_t npx_live = valid ? npx : 0u;
        const uint32_t* const inrow = &S->in[lane][0];
        const uint32_t* const inbnd = &S->in[32][0];
        uint32_t* const outrow = &S->out[lane][0];
        const bool hasb = bnd != nullptr;
        for (int c = 0; c < NSC; ++c) {
            stage(c + 2);
            cp_async_wait<1>();
            __syncwarp();
            const uint32_t cb = (uint32_t)(c * U4_CH) & (U4_INW - 1);
            const uint32_t* const ip = inrow + cb;
            const uint32_t* const bp = inbnd + cb;
            uint32_t* const op = outrow + ((uint32_t)(c * U4_CH) & (U4_OUTW - 1));
            const int x0 = c * U4_CH - lane;
            uint32_t w0 = ip[0];
#pragma unroll
            for (int s2 = 0; s2 < U4_CH; ++s2) {
                const int x = x0 + s2;
                const uint32_t upsh = __shfl_up_sync(0xffffffffu, cur, 1);
                const uint32_t bv = bp[s2];
                const uint32_t up = lane ? upsh : (hasb ? bv : 0u);
                const uint32_t w1 = s2 + 1 < U4_CH ? ip[s2 + 1] : inrow[(cb + U4_CH) & (U4_INW - 1)];
                const uint32_t raw = __funnelshift_r(w0, w1, sh);
                w0 = w1;
                // no special case for the first pixel of a row: `cur` (= the pixel to the left, and what the lane below
                // receives as its pixel above) only changes on a lane's active steps and starts the band at 0, so at
                // x == 0 both the left and the upper-left pixel read 0, as stbi__create_png_image_raw has it
                // (stbdec.d:1444-1466); one unsigned compare covers x < 0, x >= npx and rows past the image
                uint32_t pred = (cur & mS) | (up & mU);
                if (anyA) pred |= __vhaddu4(cur, up) & mA;
                if (anyP) pred |= paeth4(cur, up, upleft) & mP;
                const uint32_t nv = __vadd4(raw, pred);
                if ((uint32_t)x < npx_live) cur = nv;
                op[s2] = nv;
                upleft = up;
            }
            __syncwarp();
            flush(c);
        }
        cp_async_wait<0>();
        __syncwarp();
        if (lane == 0) flushed[warp] = (kband + 1) * (int)npx;
    }
}

// ---------------------------------------------------------------------------------------------
// Row unfilter, images whose rows are all None/Sub/Up (4-byte pixels): no pixel depends on anything but the
// pixel above (Up) or the raw bytes to its left (Sub is a prefix sum of the *filtered* row). One CTA per image,
// one warp per 128-pixel column block, every warp streams down the image with the row above in registers and
// 16-byte vector accesses; the warps of a CTA read and write whole rows together (good DRAM locality). A Sub
// row needs the sum of the blocks to its left: exchanged through shared memory with one barrier per Sub row.
constexpr int RP_MAX_WARPS = 32;
__global__ void __launch_bounds__(RP_MAX_WARPS * 32, 1)
unfilter_rowpar_kernel(const UnfilterJob* jobs, int njobs, const InflateJob* inf)
{
    __shared__ uint32_t tot[2][RP_MAX_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x;
    if (j >= njobs) return;
    const UnfilterJob J = jobs[j];
    if (inf && J.inflate_idx >= 0) {
        const InflateJob& ij = inf[J.inflate_idx];
        if (ij.status != INF_OK || ij.out_len < J.need_len) return;     // reported by the general kernel
    }
    const uint32_t rb = J.row_bytes, H = J.height, npx = rb >> 2;
    const bool eligible = J.bpp == 4 && (rb & 3) == 0 && rb >= 4 && (J.out_pitch & 3) == 0 && (((uintptr_t)J.out) & 3) == 0 &&
                          (J.inflate_idx >= 0 || (((uintptr_t)J.raw) & 15) == 0) && (npx + 127) / 128 <= blockDim.x / 32;
    if (!eligible) return;
    int wave = 0;
    for (uint32_t r = threadIdx.x; r < H; r += blockDim.x) wave |= J.raw[(size_t)r * (rb + 1)] > 2;
    if (__syncthreads_or(wave)) return;                     // has Avg/Paeth (or an invalid filter): general kernel
    const uint32_t px = (uint32_t)warp * 128 + lane * 4;
    const bool have = (uint32_t)warp * 128 < npx;           // warp-uniform: this warp owns a column block
    const uint32_t pxl = px < npx ? px : 0;                 // clamped for loads
    const bool out16 = ((J.out_pitch & 15) == 0) && ((((uintptr_t)J.out) & 15) == 0);
    const uint8_t* rp = J.raw + 1 + (size_t)pxl * 4;
    uint8_t* o = J.out + (size_t)px * 4;
    uint32_t up0 = 0, up1 = 0, up2 = 0, up3 = 0;
    int nsub = 0;
    for (uint32_t y0 = 0; y0 < H; y0 += 4) {
        uint4 A[4], B[4]; int F[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint8_t* q = rp + (size_t)min(y0 + u, H - 1) * (rb + 1);
            const uint4* v = (const uint4*)(q - ((uintptr_t)q & 15));
            A[u] = __ldg(v); B[u] = __ldg(v + 1);
            F[u] = J.raw[(size_t)min(y0 + u, H - 1) * (rb + 1)];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t y = y0 + u;
            if (y >= H) break;                              // uniform
            const uint8_t* q = rp + (size_t)y * (rb + 1);
            const uint32_t m = (uint32_t)(uintptr_t)q & 15u, shb = (m & 3u) * 8u;
            uint32_t w0, w1, w2, w3, w4;
            switch (m >> 2) {
            case 0: w0 = A[u].x; w1 = A[u].y; w2 = A[u].z; w3 = A[u].w; w4 = B[u].x; break;
            case 1: w0 = A[u].y; w1 = A[u].z; w2 = A[u].w; w3 = B[u].x; w4 = B[u].y; break;
            case 2: w0 = A[u].z; w1 = A[u].w; w2 = B[u].x; w3 = B[u].y; w4 = B[u].z; break;
            default: w0 = A[u].w; w1 = B[u].x; w2 = B[u].y; w3 = B[u].z; w4 = B[u].w; break;
            }
            uint32_t v0 = __funnelshift_r(w0, w1, shb), v1 = __funnelshift_r(w1, w2, shb);
            uint32_t v2 = __funnelshift_r(w2, w3, shb), v3 = __funnelshift_r(w3, w4, shb);
            if (px >= npx) v0 = 0; if (px + 1 >= npx) v1 = 0; if (px + 2 >= npx) v2 = 0; if (px + 3 >= npx) v3 = 0;
            if (F[u] == 1) {                                // Sub: CTA-wide prefix sum of the filtered row
                v1 = __vadd4(v1, v0); v2 = __vadd4(v2, v1); v3 = __vadd4(v3, v2);
                uint32_t inc = v3;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc = __vadd4(inc, n); }
                uint32_t ex = __shfl_up_sync(0xffffffffu, inc, 1);
                if (lane == 0) ex = 0;
                uint32_t* t = tot[nsub & 1];
                if (lane == 31) t[warp] = inc;
                __syncthreads();
                uint32_t c = 0;
                for (int b = 0; b < warp; ++b) c = __vadd4(c, t[b]);
                ex = __vadd4(ex, c);
                v0 = __vadd4(v0, ex); v1 = __vadd4(v1, ex); v2 = __vadd4(v2, ex); v3 = __vadd4(v3, ex);
                ++nsub;
            } else if (F[u] == 2) {
                v0 = __vadd4(v0, up0); v1 = __vadd4(v1, up1); v2 = __vadd4(v2, up2); v3 = __vadd4(v3, up3);
            }
            up0 = v0; up1 = v1; up2 = v2; up3 = v3;
            if (have) {
                uint32_t* op = (uint32_t*)(o + (size_t)y * J.out_pitch);
                if (px + 3 < npx && out16) __stcs((uint4*)op, make_uint4(v0, v1, v2, v3));
                else { if (px < npx) op[0] = v0; if (px + 1 < npx) op[1] = v1; if (px + 2 < npx) op[2] = v2; if (px + 3 < npx) op[3] = v3; }
            }
        }
    }
}

// Two launches per batch: <true> handles the images whose rows are all None/Sub/Up (row-parallel mode only,
// no shared-memory rings => 2x the resident warps), <false> handles every other image. Each CTA classifies its
// image by scanning the filter bytes first.
template <bool ROWPAR_ONLY>
__global__ void __launch_bounds__(U4_NW * 32, ROWPAR_ONLY ? 6 : 4)
unfilter_kernel(const UnfilterJob* jobs, int njobs, int* status, const InflateJob* inf, int rowpar_warps)
{
    extern __shared__ __align__(16) uint8_t u4_smem[];
    __shared__ volatile int flushed[U4_NW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x;
    if (j >= njobs) return;
    UnfilterJob J = jobs[j];
    if (inf && J.inflate_idx >= 0) {
        // the inflated stream must be complete and long enough, else the image fails as a whole
        const InflateJob& ij = inf[J.inflate_idx];
        if (ij.status != INF_OK || ij.out_len < J.need_len) { if (threadIdx.x == 0 && !ROWPAR_ONLY) status[J.image] = 0; return; }
    }
    const bool fast = J.bpp == 4 && (J.row_bytes & 3) == 0 && J.row_bytes >= 4 && (J.out_pitch & 3) == 0 &&
                      (((uintptr_t)J.out) & 3) == 0 &&
                      // 16-byte staging reads up to 31 bytes around each row: fine inside the decoder's padded
                      // buffers; the stand-alone entry point must hand in 16-byte aligned, padded streams
                      (J.inflate_idx >= 0 || (((uintptr_t)J.raw) & 15) == 0);
    if (fast) {
        int wave = 0;
        for (uint32_t r = threadIdx.x; r < J.height; r += U4_NW * 32) wave |= J.raw[(size_t)r * (J.row_bytes + 1)] > 2;
        wave = __syncthreads_or(wave);
        if (ROWPAR_ONLY == (wave != 0) && ((int)((J.row_bytes / 4 + 127) / 128) <= rowpar_warps)) return;   // the other launch owns this image
        if (threadIdx.x < U4_NW) flushed[threadIdx.x] = 0;
        __syncthreads();
        unfilter4_cta(J, status, ROWPAR_ONLY ? nullptr : (U4Smem*)u4_smem + warp, flushed, warp, lane);
        return;
    }
    if (ROWPAR_ONLY || warp != 0) return;
    switch (J.bpp) {
    case 1: unfilter_warp<1>(J, status, lane); break;
    case 2: unfilter_warp<2>(J, status, lane); break;
    case 3: unfilter_warp<3>(J, status, lane); break;
    case 4: unfilter_warp<4>(J, status, lane); break;
    case 6: unfilter_warp<6>(J, status, lane); break;
    default: unfilter_warp<8>(J, status, lane); break;
    }
}
void launch_unfilter(const UnfilterJob* d_jobs, int njobs, int* d_status, const InflateJob* d_inf, cudaStream_t st,
                     int rowpar_threads)
{
    if (njobs <= 0) return;
    if (rowpar_threads < 32) rowpar_threads = 32;
    if (rowpar_threads > RP_MAX_WARPS * 32) rowpar_threads = RP_MAX_WARPS * 32;
    // function attributes are per device: one flag per device, set once each (a process may switch devices)
    static std::atomic<unsigned long long> attr_mask{0};
    const int smem = (int)sizeof(U4Smem) * U4_NW;
    const int dev = device_index();
    const unsigned long long bit = dev >= 0 && dev < 64 ? 1ull << dev : 0;
    if (!bit || !(attr_mask.load(std::memory_order_acquire) & bit)) {
        if (cudaFuncSetAttribute(unfilter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess)
            attr_mask.fetch_or(bit, std::memory_order_release);
        else cudaGetLastError();
    }
    unfilter_rowpar_kernel<<<njobs, rowpar_threads, 0, st>>>(d_jobs, njobs, d_inf);
    unfilter_kernel<false><<<njobs, U4_NW * 32, smem, st>>>(d_jobs, njobs, d_status, d_inf, rowpar_threads / 32);
    count_launch(2);
}

// ---------------------------------------------------------------------------------------------
// Finish: one thread per output pixel; every remaining step of the reference pipeline, in source order.
__device__ __forceinline__ int compute_y(int r, int g, int b) { return ((r * 77) + (g * 150) + (29 * b)) >> 8; }   // stbdec.d:911

__global__ void __launch_bounds__(256)
png_finish_kernel(const FinishJob* jobs)
{
    const FinishJob& J = jobs[blockIdx.y];
    const uint32_t W = J.w, H = J.h;
    const uint64_t npix = (uint64_t)W * H;
    const int depth = J.depth, img_n = J.img_n;
    const int maxv = depth == 16 ? 65535 : 255;
    const int scale = (J.color == 0) ? (depth == 1 ? 0xff : depth == 2 ? 0x55 : depth == 4 ? 0x11 : 1) : 1;   // stbdec.d:1403,1560
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < npix; i += (uint64_t)gridDim.x * 256) {
        uint32_t y = (uint32_t)(i / W), x = (uint32_t)(i - (uint64_t)y * W);
        int p = 0; uint32_t px = x, py = y;
        if (J.interlace) {
            // Adam7 (stbdec.d:1650-1653): find the pass that owns (x, y)
            const int xorig[7] = {0, 4, 0, 2, 0, 1, 0}, yorig[7] = {0, 0, 4, 0, 2, 0, 1};
            const int xspc[7] = {8, 8, 4, 4, 2, 2, 1}, yspc[7] = {8, 8, 8, 4, 4, 2, 2};
#pragma unroll
            for (int q = 6; q >= 0; --q) {
                if ((int)x >= xorig[q] && (int)y >= yorig[q] && (x - xorig[q]) % xspc[q] == 0 && (y - yorig[q]) % yspc[q] == 0) {
                    p = q; px = (x - xorig[q]) / xspc[q]; py = (y - yorig[q]) / yspc[q];
                }
            }
        }
        const uint8_t* row = J.packed + J.pass_off[p] + (size_t)py * J.pass_rb[p];
        int s[4] = {0, 0, 0, 0};
        int n = img_n;
        if (depth == 8) {
            for (int k = 0; k < img_n; ++k) s[k] = row[(size_t)px * img_n + k];
        } else if (depth == 16) {
            for (int k = 0; k < img_n; ++k) { const uint8_t* q = row + ((size_t)px * img_n + k) * 2; s[k] = (q[0] << 8) | q[1]; }
        } else {
            for (int k = 0; k < img_n; ++k) {
                uint32_t idx = px * img_n + k;
                uint32_t bit = idx * depth;
                int byte = row[bit >> 3];
                int v = (byte >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
                s[k] = (scale * v) & 255;
            }
        }
        if (J.add_alpha) s[n++] = maxv;
        if (J.has_trans) {
            if (depth == 16) {
                if (n == 2) s[1] = (s[0] == J.tc16[0]) ? 0 : 65535;
                else if (s[0] == J.tc16[0] && s[1] == J.tc16[1] && s[2] == J.tc16[2]) s[3] = 0;
            } else {
                if (n == 2) s[1] = (s[0] == J.tc[0]) ? 0 : 255;
                else if (s[0] == J.tc[0] && s[1] == J.tc[1] && s[2] == J.tc[2]) s[3] = 0;
            }
        }
        if (J.pal_n) {
            int idx = s[0] * 4;
            for (int k = 0; k < 4; ++k) s[k] = J.palette[idx + k];
            n = J.pal_n;
        }
        const int req = J.req_n;
        int o[4] = {s[0], s[1], s[2], s[3]};
        if (req != n) {
            // stbi__convert_format / stbi__convert_format16 (stbdec.d:916-1200)
            switch (n * 8 + req) {
            case 1*8+2: o[1] = maxv; break;
            case 1*8+3: o[1] = o[2] = s[0]; break;
            case 1*8+4: o[1] = o[2] = s[0]; o[3] = maxv; break;
            case 2*8+1: break;
            case 2*8+3: o[1] = o[2] = s[0]; break;
            case 2*8+4: o[1] = o[2] = s[0]; o[3] = s[1]; break;
            case 3*8+4: o[3] = maxv; break;
            case 3*8+1: o[0] = compute_y(s[0], s[1], s[2]) & maxv; break;
            case 3*8+2: o[0] = compute_y(s[0], s[1], s[2]) & maxv; o[1] = maxv; break;
            case 4*8+1: o[0] = compute_y(s[0], s[1], s[2]) & maxv; break;
            case 4*8+2: o[0] = compute_y(s[0], s[1], s[2]) & maxv; o[1] = s[3]; break;
            case 4*8+3: break;
            }
        }
        if (J.out16) {
            uint16_t* dst = (uint16_t*)J.out + i * req;
            for (int k = 0; k < req; ++k) dst[k] = (uint16_t)(depth == 16 ? o[k] : (o[k] << 8) + o[k]);       // stbdec.d:662
        } else {
            uint8_t* dst = J.out + i * req;
            for (int k = 0; k < req; ++k) dst[k] = (uint8_t)(depth == 16 ? (o[k] >> 8) & 0xFF : o[k]);         // stbdec.d:645
        }
    }
}
void launch_finish(const FinishJob* d_jobs, int njobs, uint64_t max_pixels, cudaStream_t st)
{
    if (njobs <= 0) return;
    uint64_t bx = (max_pixels + 255) / 256;
    if (bx > 148 * 32) bx = 148 * 32;
    if (bx < 1) bx = 1;
    for (int base = 0; base < njobs; base += 65535) {
        int ny = njobs - base < 65535 ? njobs - base : 65535;
        png_finish_kernel<<<dim3((unsigned)bx, (unsigned)ny), 256, 0, st>>>(d_jobs + base);
        count_launch();
    }
}

} // namespace gb
