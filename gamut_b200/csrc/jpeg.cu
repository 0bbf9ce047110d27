// jpeg.cu -- JPEG decode (sequential and progressive Huffman) for sm_100a.
//
// Drop-in for decompress_jpeg_image_from_stream (source/gamut/codecs/jpegload.d:3720-3808) as used by
// loadJPEG (source/gamut/plugins/jpeg.d:42-104). The marker layer (jpegload.d:1177-1967: DQT, DHT, SOF,
// DRI, SOS, JFIF/EXIF density) is header logic and runs on the host; every per-bit / per-coefficient /
// per-pixel step runs on the GPU:
//   entropy stage         long segments: jpeg_unstuff_* / jpeg_sync / jpeg_repair / jpeg_scan / jpeg_write
//                         (jpeg_sync.cuh: chunk-parallel, self-synchronising); segments under 1 KB:
//                         jpeg_huffman_kernel, one thread per segment; progressive files: jpeg_prog_kernel +
//                         jpeg_prog_gather_kernel. Entropy decode + dequantisation into natural-order int16
//                         blocks (decode_next_row, jpegload.d:2405-2525; huff_decode :746-813)
//   jpeg_idct_colour_kernel  integer LL&M IDCT (Row/Col, :156-397) and, for 4:2:0, the frequency-domain chroma
//                         upsampling DCT_Upsample + idct_4x4 (:827-1073, :2139-2255), fused with YCbCr ->
//                         RGB(A) / Y with the 16.16 fixed-point constants of create_look_ups (:2080-2094,
//                         H*Convert :2528-2823) and the final channel adaptation (:3763-3801): the sample
//                         tiles never leave shared memory
// Integer arithmetic throughout; results are bit-exact with the restated reference.
#include "common.h"
#include "batch.h"
#include <vector>
#include <map>
#include <string>
#include <chrono>
#include <cmath>

namespace {

enum { GRAYSCALE = 0, YH1V1, YH2V1, YH1V2, YH2V2 };

// ---- device-side structures ------------------------------------------------------------------
constexpr int HUFF_FAST = 10;
constexpr int HT_L1_BITS = 9, HT_L2_BITS = 7, HT_SUBS = 6;
static_assert(true, "");
constexpr uint16_t HT_LONG = 0x8000;
constexpr int HT_KINC_EOB = 63;      // zig-zag advance that ends the block from any AC position (k >= 1)
struct alignas(16) HuffTable {   // canonical JPEG code (Annex C) + 10-bit lookahead
    uint16_t fast[1 << HUFF_FAST];   // (len << 8) | symbol, 0 = longer than HUFF_FAST bits / invalid
    int32_t maxcode[18];             // maxcode[l] for l = 1..16, -1 if none
    int32_t mincode[18];
    int32_t valptr[18];
    uint8_t val[256];
    // two-level table of the chunk-parallel decoders (copied to shared memory): entry = len (bits 0-4, 0 = not a code
    // word) | size << 5 | kinc << 9, kinc = how far the symbol moves the zig-zag index: 1 for a DC symbol, run + 1 for
    // an AC coefficient, 16 for ZRL, HT_KINC_EOB for EOB (jpegload.d:2466-2507) -- one add per symbol instead of a
    // case distinction. An l1 entry with HT_LONG set points at the l2 sub-table (bits 0-2) indexed by the next 7 bits;
    // sub-table 7 = more long-code prefixes than sub-tables (decode through maxcode/valptr).
    int32_t is_dc;
    alignas(16) uint16_t l1[1 << HT_L1_BITS];     // 16-byte aligned: copied to shared memory by 16-byte vectors
    uint16_t l2[HT_SUBS][1 << HT_L2_BITS];
};

struct JpegImage {
    const uint8_t* data;         // device copy of the file
    uint32_t data_len;
    int width, height;
    int scan_type, comps;        // comps_in_frame: 1 or 3
    int mcus_per_row, mcus_per_col, blocks_per_mcu, tiles_per_mcu;
    int mcu_org[10];             // component of each block in the MCU
    int dc_tab[3], ac_tab[3];    // indices into the global HuffTable array
    int16_t quant[3][64];        // per component, zig-zag order (jpegload.d:1312-1327)
    int16_t* coefs;              // [mcu][block][64]
    uint8_t* blk_zag;            // [mcu][block]: zig-zag extent of the block (index of its last non-zero coefficient + 1)
    uint8_t* samples;            // [mcu][tile][64]
    uint8_t* out;
    int req_comps;
    int restart_interval;
};

struct Segment {                 // one independently decodable run of MCUs
    int image;
    uint32_t start, end;         // byte range of the entropy data (end = first byte that is not data)
    int first_mcu, num_mcus;
};

__constant__ uint8_t c_zag[64] = {0,1,8,16,9,2,3,10,17,24,32,25,18,11,4,5,12,19,26,33,40,48,41,34,27,20,13,6,7,14,21,28,
                                  35,42,49,56,57,50,43,36,29,22,15,23,30,37,44,51,58,59,52,45,38,31,39,46,53,60,61,54,47,55,62,63};

// ---- entropy decode ---------------------------------------------------------------------------
// Bit source with the semantics of get_bits_no_markers/get_octet (jpegload.d:683-743): FF 00 is a
// literal FF, any other FF xx (a marker) and the end of the data yield 1-bits forever.
struct BitReader {
    const uint8_t* p; uint32_t pos, end;
    uint64_t buf; int cnt;       // `cnt` valid bits, MSB-aligned in buf
    bool ended;
    __device__ __forceinline__ void init(const uint8_t* data, uint32_t s, uint32_t e) { p = data; pos = s; end = e; buf = 0; cnt = 0; ended = false; }
    __device__ __forceinline__ void fill()
    {
        while (cnt <= 56) {
            uint32_t b = 0xFF;
            if (!ended) {
                if (pos >= end) ended = true;
                else {
                    b = p[pos];
                    if (b == 0xFF) {
                        if (pos + 1 < end && p[pos + 1] == 0) pos += 2;
                        else { ended = true; }
                    } else pos += 1;
                }
            }
            buf |= (uint64_t)b << (56 - cnt);
            cnt += 8;
        }
    }
    __device__ __forceinline__ uint32_t peek16() { return (uint32_t)(buf >> 48); }
    __device__ __forceinline__ void drop(int n) { buf <<= n; cnt -= n; }
    __device__ __forceinline__ uint32_t get(int n) { if (n == 0) return 0; uint32_t v = (uint32_t)(buf >> (64 - n)); drop(n); return v; }
};

__device__ __forceinline__ int huff_decode(BitReader& br, const HuffTable* __restrict__ h)
{
    br.fill();
    uint32_t top = br.peek16();
    uint32_t e = h->fast[top >> (16 - HUFF_FAST)];
    if (e) { br.drop(e >> 8); return e & 255; }
    for (int l = HUFF_FAST + 1; l <= 16; ++l) {
        int code = (int)(top >> (16 - l));
        if (code <= h->maxcode[l] && code >= h->mincode[l]) { br.drop(l); return h->val[h->valptr[l] + code - h->mincode[l]]; }
    }
    // codes of length <= HUFF_FAST not present in the fast table are invalid as well
    return -1;
}
__device__ __forceinline__ int huff_extend(int x, int s) { return (s == 0) ? x : ((x < (1 << (s - 1))) ? x + (int)(0xFFFFFFFFu << s) + 1 : x); }

__global__ void __launch_bounds__(128)
jpeg_huffman_kernel(const JpegImage* __restrict__ images, const Segment* __restrict__ segs, int nsegs,
                    const HuffTable* __restrict__ tables, int* status)
{
    int si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= nsegs) return;
    const Segment sg = segs[si];
    const JpegImage& im = images[sg.image];
    BitReader br; br.init(im.data, sg.start, sg.end);
    int dc[3] = {0, 0, 0};
    const int bpm = im.blocks_per_mcu;
    int16_t* coef = im.coefs + (size_t)sg.first_mcu * bpm * 64;
    uint8_t* zag = im.blk_zag + (size_t)sg.first_mcu * bpm;
    for (int m = 0; m < sg.num_mcus; ++m) {
        for (int b = 0; b < bpm; ++b, coef += 64, ++zag) {
            int last_k = 0;
            const int comp = im.mcu_org[b];
            const int16_t* q = im.quant[comp];
            int s = huff_decode(br, tables + im.dc_tab[comp]);
            if (s < 0 || s > 15) { status[sg.image] = 0; return; }
            br.fill();
            int r = (int)br.get(s);
            s = huff_extend(r, s);
            dc[comp] = (s += dc[comp]);
            coef[0] = (int16_t)(s * q[0]);
            const HuffTable* ac = tables + im.ac_tab[comp];
            for (int k = 1; k < 64; ++k) {
                int rs = huff_decode(br, ac);
                if (rs < 0) { status[sg.image] = 0; return; }
                br.fill();
                int extra = (int)br.get(rs & 15);
                r = rs >> 4; s = rs & 15;
                if (s) {
                    if (r) { if (k + r > 63) { status[sg.image] = 0; return; } k += r; }
                    s = huff_extend(extra, s);
                    coef[c_zag[k]] = (int16_t)(s * q[k]);
                    last_k = k;
                } else {
                    if (r == 15) { if (k + 16 > 64) { status[sg.image] = 0; return; } k += 15; }
                    else break;
                }
            }
            *zag = (uint8_t)(last_k + 1);
        }
    }
}

// ---- progressive (SOF2) ---------------------------------------------------------------------------------------------
// init_progressive / decode_scan / decode_block_* (jpegload.d:3299-3684): every scan of the file refines whole-image
// coefficient planes (one per component, DC in element 0 of each 64-coefficient block: the reference keeps DC and AC
// in two buffers and joins them in load_next_row, :2280-2284); scans must run in file order (a refinement scan reads
// what earlier scans wrote), and a scan is one serial bit stream, so one thread walks all the scans of one image.
// Parallelism is across the images of a batch; the dequantisation, IDCT and colour stages are the batch-parallel
// kernels of the sequential path (jpeg_prog_gather_kernel puts the planes into the MCU order they read).
struct ProgImage {
    int image;                 // index into JpegImage[]
    int16_t* plane[3];         // [block_y][block_x][64]
    int num_x[3], num_y[3], h[3], v[3];
    int first_scan, nscans;
};
struct ProgScan {
    int ncomp, comp[3], dc_tab[3], ac_tab[3];      // indices into the global HuffTable array
    int ss, se, ah, al;
    int mcus_per_row;
    int first_seg, nsegs;      // Segment[]: one per restart interval (image = index of the ProgImage)
};

// one block of a scan; kind = (AC scan ? 2 : 0) + (refinement ? 1 : 0). Returns false on a decode error.
__device__ __forceinline__ bool prog_block(BitReader& br, const HuffTable* __restrict__ dct, const HuffTable* __restrict__ act,
                                           const ProgScan& sc, int kind, int16_t* __restrict__ p, int& last_dc, int& eob_run)
{
    const int al = sc.al;
    if (kind == 0) {                                    // decode_block_dc_first (:3299-3320)
        int s = huff_decode(br, dct);
        if (s < 0 || s > 15) return false;
        if (s != 0) { br.fill(); const int r = (int)br.get(s); s = huff_extend(r, s); }
        last_dc = (s += last_dc);
        p[0] = (int16_t)((uint32_t)s << al);
        return true;
    }
    if (kind == 1) {                                    // decode_block_dc_refine (:3322-3333)
        br.fill();
        if (br.get(1)) p[0] = (int16_t)(p[0] | (1 << al));
        return true;
    }
    if (kind == 2) {                                    // decode_block_ac_first (:3335-3398)
        if (eob_run) { --eob_run; return true; }
        for (int k = sc.ss; k <= sc.se; ++k) {
            const int rs = huff_decode(br, act);
            if (rs < 0) return false;
            int r = rs >> 4, s = rs & 15;
            if (s) {
                if ((k += r) > 63) return false;
                br.fill();
                r = (int)br.get(s);
                s = huff_extend(r, s);
                p[c_zag[k]] = (int16_t)((uint32_t)s << al);
            } else if (r == 15) {
                if ((k += 15) > 63) return false;
            } else {
                eob_run = 1 << r;
                if (r) { br.fill(); eob_run += (int)br.get(r); }
                --eob_run;
                break;
            }
        }
        return true;
    }
    // decode_block_ac_refine (:3400-3519)
    const int p1 = 1 << al, m1 = (int)(0xFFFFFFFFu << al);
    int k = sc.ss;
    if (eob_run == 0) {
        for (; k <= sc.se; ++k) {
            const int rs = huff_decode(br, act);
            if (rs < 0) return false;
            int r = rs >> 4, s = rs & 15;
            if (s) {
                if (s != 1) return false;
                br.fill();
                s = br.get(1) ? p1 : m1;
            } else if (r != 15) {
                eob_run = 1 << r;
                if (r) { br.fill(); eob_run += (int)br.get(r); }
                break;
            }
            do {
                int16_t* tc = p + c_zag[k & 63];
                if (*tc != 0) {
                    br.fill();
                    if (br.get(1)) { if ((*tc & p1) == 0) *tc = (int16_t)(*tc + (*tc >= 0 ? p1 : m1)); }
                } else if (--r < 0) break;
                ++k;
            } while (k <= sc.se);
            if (s && k < 64) p[c_zag[k]] = (int16_t)s;
        }
    }
    if (eob_run > 0) {
        for (; k <= sc.se; ++k) {
            int16_t* tc = p + c_zag[k & 63];
            if (*tc != 0) {
                br.fill();
                if (br.get(1)) { if ((*tc & p1) == 0) *tc = (int16_t)(*tc + (*tc >= 0 ? p1 : m1)); }
            }
        }
        --eob_run;
    }
    return true;
}

__global__ void __launch_bounds__(32)
jpeg_prog_kernel(const JpegImage* __restrict__ images, const ProgImage* __restrict__ prog, int nprog,
                 const ProgScan* __restrict__ scans, const Segment* __restrict__ segs,
                 const HuffTable* __restrict__ tables, int* status)
{
    const int pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= nprog) return;
    const ProgImage& P = prog[pi];
    if (!status[P.image]) return;
    const uint8_t* data = images[P.image].data;
    for (int si = 0; si < P.nscans; ++si) {
        const ProgScan& sc = scans[P.first_scan + si];
        const int kind = (sc.ss ? 2 : 0) + (sc.ah ? 1 : 0);
        for (int g = 0; g < sc.nsegs; ++g) {
            const Segment sg = segs[sc.first_seg + g];
            BitReader br; br.init(data, sg.start, sg.end);
            int dc[3] = {0, 0, 0};
            int eob_run = 0;                            // process_restart clears it with the DC predictors (:2380-2386)
            for (int m = sg.first_mcu; m < sg.first_mcu + sg.num_mcus; ++m) {
                const int mx = m % sc.mcus_per_row, my = m / sc.mcus_per_row;
                for (int i = 0; i < sc.ncomp; ++i) {
                    const int c = sc.comp[i];
                    const int hh = sc.ncomp == 1 ? 1 : P.h[c], vv = sc.ncomp == 1 ? 1 : P.v[c];
                    const HuffTable* dct = tables + sc.dc_tab[i];
                    const HuffTable* act = tables + sc.ac_tab[i];
                    for (int yo = 0; yo < vv; ++yo) {
                        for (int xo = 0; xo < hh; ++xo) {
                            const int bx = mx * hh + xo, by = my * vv + yo;
                            if (bx >= P.num_x[c] || by >= P.num_y[c]) { status[P.image] = 0; return; }     // the reference asserts
                            int16_t* p = P.plane[c] + ((size_t)by * P.num_x[c] + bx) * 64;
                            if (!prog_block(br, dct, act, sc, kind, p, dc[c], eob_run)) { status[P.image] = 0; return; }
                        }
                    }
                }
            }
        }
    }
}

// load_next_row (jpegload.d:2259-2332) for the whole image: one warp per block of the interleaved MCU grid copies the
// block out of its component plane, dequantises it (quant tables in zig-zag order, products kept as 16-bit like the
// reference's cast) and records the zig-zag extent the IDCT stage uses to pick its sparse variant.
__global__ void __launch_bounds__(128)
jpeg_prog_gather_kernel(const JpegImage* __restrict__ images, const ProgImage* __restrict__ prog, const uint32_t* __restrict__ blk_base,
                        int nprog, const int* __restrict__ status)
{
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= blk_base[nprog]) return;
    int lo = 0, hi = nprog - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (blk_base[mid] <= gw) lo = mid; else hi = mid - 1; }
    const ProgImage& P = prog[lo];
    if (!status[P.image]) return;
    const JpegImage& im = images[P.image];
    const uint32_t lb = gw - blk_base[lo];
    const int bpm = im.blocks_per_mcu;
    const int m = (int)(lb / (uint32_t)bpm), b = (int)(lb - (uint32_t)m * bpm);
    const int c = im.mcu_org[b];
    int j = 0;
    for (int t = 0; t < b; ++t) j += im.mcu_org[t] == c ? 1 : 0;
    const int mx = m % im.mcus_per_row, my = m / im.mcus_per_row;
    const int bx = mx * P.h[c] + j % P.h[c], by = my * P.v[c] + j / P.h[c];
    const int16_t* __restrict__ src = P.plane[c] + ((size_t)by * P.num_x[c] + bx) * 64;
    const int k0 = lane, k1 = lane + 32;
    const int n0 = c_zag[k0], n1 = c_zag[k1];
    const int v0 = src[n0], v1 = src[n1];
    const uint32_t z0 = __ballot_sync(0xffffffffu, v0 != 0) & ~1u, z1 = __ballot_sync(0xffffffffu, v1 != 0);
    const int last = z1 ? 63 - __clz(z1) : (z0 ? 31 - __clz(z0) : 0);
    int16_t* __restrict__ dst = im.coefs + ((size_t)m * bpm + b) * 64;
    dst[n0] = (int16_t)(v0 * im.quant[c][k0]);
    dst[n1] = (int16_t)(v1 * im.quant[c][k1]);
    if (lane == 0) im.blk_zag[(size_t)m * bpm + b] = (uint8_t)(last + 1);
}

#include "jpeg_sync.cuh"

// ---- IDCT -------------------------------------------------------------------------------------
#define CONST_BITS 13
#define PASS1_BITS 2
#define FIX_0_298631336 2446
#define FIX_0_390180644 3196
#define FIX_0_541196100 4433
#define FIX_0_765366865 6270
#define FIX_0_899976223 7373
#define FIX_1_175875602 9633
#define FIX_1_501321110 12299
#define FIX_1_847759065 15137
#define FIX_1_961570560 16069
#define FIX_2_053119869 16819
#define FIX_2_562915447 20995
#define FIX_3_072711026 25172

__device__ __forceinline__ int clamp255(int i) { return __vimin_s32_relu(i, 255); }     // one VIMNMX: max(min(i, 255), 0)

// One 8-point LL&M pass (jpegload.d:172-213 / :236-289). in[] are the 8 inputs; FINAL selects the
// column-pass descale (+128 level shift, >> 18, clamp) versus the row-pass descale (>> 11).
template <bool FINAL>
__device__ __forceinline__ void idct8(const int in[8], int out[8]);
// N leading inputs may be non-zero (Row!N / Col!N of the reference, jpegload.d:156-292): the rest are literal zeros
template <bool FINAL, int N>
__device__ __forceinline__ void idct8n(const int src[8], int out[8])
{
    int in[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) in[i] = i < N ? src[i] : 0;
    idct8<FINAL>(in, out);
}
template <bool FINAL>
__device__ __forceinline__ void idct8(const int in[8], int out[8])
{
    const int z2 = in[2], z3 = in[6];
    const int z1 = (z2 + z3) * FIX_0_541196100;
    const int tmp2 = z1 + z3 * (-FIX_1_847759065);
    const int tmp3 = z1 + z2 * FIX_0_765366865;
    // The descale rounding term (and the +128 level shift of the final pass) is added to the even part once instead of
    // to each of the eight outputs: int32 addition is associative modulo 2^32, so every output is the same integer as
    // DESCALE / DESCALE_ZEROSHIFT of the reference (jpegload.d:137-145).
    constexpr int RND = FINAL ? (128 << (CONST_BITS + PASS1_BITS + 3)) + (1 << (CONST_BITS + PASS1_BITS + 2)) : (1 << (CONST_BITS - PASS1_BITS - 1));
    constexpr int SH = FINAL ? CONST_BITS + PASS1_BITS + 3 : CONST_BITS - PASS1_BITS;
    const int tmp0 = (int)((unsigned)(in[0] + in[4]) * (1u << CONST_BITS) + (unsigned)RND);     // unsigned: wraps like the reference's int
    const int tmp1 = (int)((unsigned)(in[0] - in[4]) * (1u << CONST_BITS) + (unsigned)RND);
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    const int atmp0 = in[7], atmp1 = in[5], atmp2 = in[3], atmp3 = in[1];
    const int bz1 = atmp0 + atmp3, bz2 = atmp1 + atmp2, bz3 = atmp0 + atmp2, bz4 = atmp1 + atmp3;
    const int bz5 = (bz3 + bz4) * FIX_1_175875602;
    const int az1 = bz1 * (-FIX_0_899976223);
    const int az2 = bz2 * (-FIX_2_562915447);
    const int az3 = bz3 * (-FIX_1_961570560) + bz5;
    const int az4 = bz4 * (-FIX_0_390180644) + bz5;
    const int btmp0 = atmp0 * FIX_0_298631336 + az1 + az3;
    const int btmp1 = atmp1 * FIX_2_053119869 + az2 + az4;
    const int btmp2 = atmp2 * FIX_3_072711026 + az2 + az3;
    const int btmp3 = atmp3 * FIX_1_501321110 + az1 + az4;
    const int s[8] = {tmp10 + btmp3, tmp11 + btmp2, tmp12 + btmp1, tmp13 + btmp0, tmp13 - btmp0, tmp12 - btmp1, tmp11 - btmp2, tmp10 - btmp3};
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = FINAL ? clamp255(s[i] >> SH) : (s[i] >> SH);
}

// DCT_Upsample (jpegload.d:827-1073): a chroma block -> four 4x4 frequency tiles -> idct_4x4 each.
__device__ __forceinline__ int UD(int i) { return (i + 512) >> 10; }


// ---- fused IDCT + chroma upsampling + colour conversion -----------------------------------------
// One CTA converts a run of consecutive MCUs of one MCU row: 8-thread groups run the two IDCT passes of a
// block (thread = row, then thread = column, transposing through shared memory), sample tiles stay in shared
// memory in the reference's m_pSample_buf tile order, and the colour pass writes whole output rows of the run
// with 16-byte stores. Coefficients are read once (16 B per thread, 128 B per group) and pixels written once.
constexpr int IC_THREADS = 256, IC_GROUPS = 32, IC_GSTRIDE = 264, IC_MAX_TILES = 192, IC_MCU420 = 12 * 64 + 16;

__device__ __forceinline__ void unpack8(const int4 v, int in[8])
{
    in[0] = (int16_t)(v.x & 0xffff); in[1] = v.x >> 16; in[2] = (int16_t)(v.y & 0xffff); in[3] = v.y >> 16;
    in[4] = (int16_t)(v.z & 0xffff); in[5] = v.z >> 16; in[6] = (int16_t)(v.w & 0xffff); in[7] = v.w >> 16;
}

__device__ __forceinline__ void ycc_to_rgb(int Y, int cb, int cr, int& r, int& g, int& b)
{
    // FIX!(x) = (int)(x * 65536 + 0.5f) (jpegload.d:2082). The reference's tables hold F * (c - 128) [+ 32768]; the
    // "- 128" is folded into the additive constant here (same integers: F * (c - 128) + K == F * c + (K - 128 * F)).
    constexpr int F140200 = 91881, F177200 = 116130, F071414 = 46802, F034414 = 22554;
    r = clamp255(Y + ((F140200 * cr + (32768 - 128 * F140200)) >> 16));
    g = clamp255(Y + (((-F071414) * cr + (-F034414) * cb + (32768 + 128 * (F071414 + F034414))) >> 16));
    b = clamp255(Y + ((F177200 * cb + (32768 - 128 * F177200)) >> 16));
}

// 16 pixels of one row of a 4:2:0 MCU: Y from two luma tiles, Cb / Cr from the frequency-domain-upsampled tiles
// (expanded_convert, jpegload.d:2731-2823) -> RC bytes per pixel (final channel adaptation :3763-3801).
template <int RC>
__device__ __forceinline__ void colour420(const uint8_t* __restrict__ tb, uint8_t* __restrict__ d, int n)
{
    uint32_t wds[RC * 4];           // 16 pixels * RC bytes
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint2 y8 = *(const uint2*)(tb + h * 64), cb8 = *(const uint2*)(tb + 256 + h * 64), cr8 = *(const uint2*)(tb + 512 + h * 64);
        uint8_t px[8 * RC];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int Y = ((i < 4 ? y8.x : y8.y) >> ((i & 3) * 8)) & 255;
            const int cb = ((i < 4 ? cb8.x : cb8.y) >> ((i & 3) * 8)) & 255;
            const int cr = ((i < 4 ? cr8.x : cr8.y) >> ((i & 3) * 8)) & 255;
            int r, g, b; ycc_to_rgb(Y, cb, cr, r, g, b);
            if (RC == 1) px[i] = (uint8_t)((r * 19595 + g * 38470 + b * 7471 + 32768) >> 16);
            else { px[i * RC] = (uint8_t)r; px[i * RC + 1] = (uint8_t)g; px[i * RC + 2] = (uint8_t)b; if (RC == 4) px[i * 4 + 3] = 255; }
        }
#pragma unroll
        for (int q = 0; q < 2 * RC; ++q)
            wds[h * 2 * RC + q] = (uint32_t)px[q * 4] | ((uint32_t)px[q * 4 + 1] << 8) | ((uint32_t)px[q * 4 + 2] << 16) | ((uint32_t)px[q * 4 + 3] << 24);
    }
    if (n == 16 && (((uintptr_t)d) & 15) == 0) {
#pragma unroll
        for (int q = 0; q < RC; ++q) ((uint4*)d)[q] = make_uint4(wds[q * 4], wds[q * 4 + 1], wds[q * 4 + 2], wds[q * 4 + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < 16 * RC; ++i) if (i < n * RC) d[i] = (uint8_t)(wds[i >> 2] >> ((i & 3) * 8));
    }
}

__global__ void __launch_bounds__(IC_THREADS, 4)
jpeg_idct_colour_kernel(const JpegImage* __restrict__ images, const uint32_t* __restrict__ cta_base, int nimages,
                        const int* __restrict__ status)
{
    // 4:2:0: the 12 tiles of an MCU are IC_MCU420 = 784 bytes apart (768 + 16): the colour pass reads 8-byte rows of 16
    // MCUs at once, and a 768-byte stride would put them all in the same banks
    __shared__ __align__(16) uint8_t s_tiles[IC_MAX_TILES * 64 + 16 * 16];
    __shared__ __align__(16) int s_tmp[IC_GROUPS * IC_GSTRIDE];
    // grid = (largest CTA count of an image, images): no search for the image (a binary search over the CTA prefix cost
    // nine dependent global loads at the start of every CTA, 17 % of the kernel's stall samples in round 2)
    const int lo = (int)blockIdx.y;
    if (lo >= nimages || blockIdx.x >= cta_base[lo + 1] - cta_base[lo]) return;
    if (!status[lo]) return;
    const JpegImage& im = images[lo];
    const int st = im.scan_type;
    const int NM = st == YH2V2 ? 16 : 32;
    const int gpr = (im.mcus_per_row + NM - 1) / NM;
    const int local = (int)blockIdx.x;
    const int mrow = local / gpr, g0 = (local - mrow * gpr) * NM;
    const int nm = min(NM, im.mcus_per_row - g0);
    const int bpm = im.blocks_per_mcu, tpm = im.tiles_per_mcu;
    const int16_t* __restrict__ coefs = im.coefs + ((size_t)mrow * im.mcus_per_row + g0) * bpm * 64;
    const int tid = threadIdx.x, grp = tid >> 3, t = tid & 7;
    int* tmp = s_tmp + grp * IC_GSTRIDE;

    // ---- stage 1a: plain 8x8 IDCT (every block except 4:2:0 chroma). The zig-zag extent of a block bounds its
    // non-zero rows and columns (s_idct_row_table / s_idct_col_table, jpegload.d:295-306); the warp takes the sparse
    // Row!N / Col!N variant that covers its four blocks: extent 1 -> DC only, <= 10 -> 4x4, <= 21 -> 6x6, else 8x8.
    const int npm = st == YH2V2 ? 4 : bpm;
    const int nplain = nm * npm;
    const uint8_t* __restrict__ zags = im.blk_zag + ((size_t)mrow * im.mcus_per_row + g0) * bpm;
    // the coefficient row and the extent of the NEXT task are fetched while the current one is transformed (the two
    // dependent global loads at the top of an iteration were 16 % of the kernel's stall samples)
    // 4:2:0: the chroma row of this group's (single) upsampling task is requested now and used after the luma blocks
    int4 chroma_v = make_int4(0, 0, 0, 0); int chroma_zag = 0;
    if (st == YH2V2 && grp < nm * 2) {
        chroma_zag = zags[(grp >> 1) * 6 + 4 + (grp & 1)];
        chroma_v = __ldg((const int4*)(coefs + ((size_t)(grp >> 1) * 6 + 4 + (grp & 1)) * 64) + t);
    }
    int4 nv = make_int4(0, 0, 0, 0); int nzag = 0;
    auto fetch = [&](int task, int4& v, int& zag) {
        v = make_int4(0, 0, 0, 0); zag = 0;
        if (task < nplain) {
            const int m = task / npm, bi = task - m * npm;
            zag = zags[m * bpm + bi];
            v = __ldg((const int4*)(coefs + ((size_t)m * bpm + bi) * 64) + t);
        }
    };
    fetch(grp, nv, nzag);
    for (int base = 0; base < nplain; base += IC_GROUPS) {
        const int task = base + grp;
        const bool active = task < nplain;
        int m = 0, bi = 0;
        if (active) { m = task / npm; bi = task - m * npm; }
        const int4 cv = nv; const int zag = nzag;
        fetch(task + IC_GROUPS, nv, nzag);
        const int zmax = __reduce_max_sync(0xffffffffu, zag);
        uint8_t* dst = s_tiles + (st == YH2V2 ? m * IC_MCU420 + bi * 64 : (m * tpm + bi) * 64);
        if (zmax <= 1) {
            // idct with block_max_zag <= 1 (jpegload.d:312-326): all 64 samples are ((dc + 4) >> 3) + 128, clamped
            const int dcv = (int)(short)__shfl_sync(0xffffffffu, cv.x, (threadIdx.x & 31) & ~7);         // element 0 of the block: row 0 is with thread t = 0
            if (active) {
                const uint32_t v = (uint32_t)clamp255(((dcv + 4) >> 3) + 128) * 0x01010101u;
                *(uint2*)(dst + t * 8) = make_uint2(v, v);
            }
        } else {
#define IC_PLAIN(N)                                                                                          \
            do {                                                                                             \
                if (active && t < N) {                                                                       \
                    int in[8], out[8];                                                                       \
                    unpack8(cv, in);                                                                         \
                    idct8n<false, N>(in, out);                                                               \
                    *(int4*)(tmp + t * 8) = make_int4(out[0], out[1], out[2], out[3]);                       \
                    *(int4*)(tmp + t * 8 + 4) = make_int4(out[4], out[5], out[6], out[7]);                   \
                }                                                                                            \
                __syncwarp();                                                                                \
                if (active) {                                                                                \
                    int in[8], out[8];                                                                       \
                    _Pragma("unroll") for (int r = 0; r < 8; ++r) in[r] = r < N ? tmp[r * 8 + t] : 0;        \
                    idct8n<true, N>(in, out);                                                                \
                    _Pragma("unroll") for (int r = 0; r < 8; ++r) dst[r * 8 + t] = (uint8_t)out[r];          \
                }                                                                                            \
            } while (0)
            if (zmax <= 10) IC_PLAIN(4);
            else if (zmax <= 21) IC_PLAIN(6);
            else IC_PLAIN(8);
#undef IC_PLAIN
        }
        __syncwarp();
    }
    // ---- stage 1b: 4:2:0 chroma, DCT_Upsample (jpegload.d:827-1073) + idct_4x4 on the four tiles
    if (st == YH2V2) {
        const int a1[4] = {426, 810, -360, 284};
        const int a2[4] = {23, -99, 502, 887};
        const int b1[4] = {928, -325, 218, -184};
        const int b2[4] = {-75, 526, 787, -383};
        const int nup = nm * 2;
        for (int base = 0; base < nup; base += IC_GROUPS) {
            const int task = base + grp;
            const bool active = task < nup;
            const int m = task >> 1, ch = task & 1;
            const bool first = base == 0;                // nm <= 16: the only iteration; kept general
            const int czag = !active ? 0 : first ? chroma_zag : (int)zags[m * 6 + 4 + ch];
            const int4 cv4 = first ? chroma_v : (active ? __ldg((const int4*)(coefs + ((size_t)m * 6 + 4 + ch) * 64) + t) : make_int4(0, 0, 0, 0));
            if (__reduce_max_sync(0xffffffffu, czag) <= 1) {
                // P_Q!(1,1) / R_S!(1,1): only P[0][0] = DC is non-zero, the four tiles are the DC-only idct_4x4:
                // Row!4 gives DC << 2 along row 0, Col!4 then ((t + 128*32 + 16) >> 5) = ((DC + 4) >> 3) + 128
                const int dcv = __shfl_sync(0xffffffffu, cv4.x, (threadIdx.x & 31) & ~7);
                if (active) {
                    const int t4 = (int)(short)dcv;          // add_and_store keeps a short
                    const uint32_t v = (uint32_t)clamp255((((t4 << 2) + (128 << 5) + 16) >> 5)) * 0x01010101u;
                    uint8_t* dstc = s_tiles + m * IC_MCU420 + (4 + ch * 4) * 64;
#pragma unroll
                    for (int tt = 0; tt < 4; ++tt) *(uint2*)(dstc + tt * 64 + t * 8) = make_uint2(v, v);
                }
                __syncwarp();
                continue;
            }
            if (active) {   // A: thread j = source row j -> X0[0..3][j], X1[0..3][j]
                int s[8];
                unpack8(cv4, s);
                tmp[0 * 8 + t] = s[0];
                tmp[1 * 8 + t] = UD(a1[0] * s[1] + a1[1] * s[3] + a1[2] * s[5] + a1[3] * s[7]);
                tmp[2 * 8 + t] = s[4];
                tmp[3 * 8 + t] = UD(a2[0] * s[1] + a2[1] * s[3] + a2[2] * s[5] + a2[3] * s[7]);
                tmp[32 + 0 * 8 + t] = UD(b1[0] * s[1] + b1[1] * s[3] + b1[2] * s[5] + b1[3] * s[7]);
                tmp[32 + 1 * 8 + t] = s[2];
                tmp[32 + 2 * 8 + t] = UD(b2[0] * s[1] + b2[1] * s[3] + b2[2] * s[5] + b2[3] * s[7]);
                tmp[32 + 3 * 8 + t] = s[6];
            }
            __syncwarp();
            if (active) {   // B: thread (i, which): P,Q rows from X0[i], R,S rows from X1[i]
                const int i = t & 3, which = t >> 2;
                const int* x = tmp + which * 32 + i * 8;
                const int4 xa = *(const int4*)x, xb = *(const int4*)(x + 4);
                const int x1 = xa.y, x3 = xa.w, x5 = xb.y, x7 = xb.w;
                int* pq = tmp + 64 + which * 32;     // [P|R][i][c] at +0, [Q|S][i][c] at +16
                *(int4*)(pq + i * 4) = make_int4(xa.x, UD(x1 * a1[0] + x3 * a1[1] + x5 * a1[2] + x7 * a1[3]),
                                                 xb.x, UD(x1 * a2[0] + x3 * a2[1] + x5 * a2[2] + x7 * a2[3]));
                *(int4*)(pq + 16 + i * 4) = make_int4(UD(x1 * b1[0] + x3 * b1[1] + x5 * b1[2] + x7 * b1[3]), xa.z,
                                                      UD(x1 * b2[0] + x3 * b2[1] + x5 * b2[2] + x7 * b2[3]), xb.z);
            }
            __syncwarp();
            if (active) {   // C: row pass of the four transposed 4x4 tiles (jpegload.d:886-902, :2230-2251)
                const int tt = t >> 1;
                const int* Pm = tmp + 64, *Qm = tmp + 80, *Rm = tmp + 96, *Sm = tmp + 112;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int rr = (t & 1) * 2 + q;     // row of the transposed tile = column c of P..S
                    int in[8], out[8];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {    // column of the transposed tile = row r of P..S
                        const int idx = cc * 4 + rr;
                        const int a = Pm[idx] + Qm[idx], b = Pm[idx] - Qm[idx], c = Rm[idx] + Sm[idx], d = Rm[idx] - Sm[idx];
                        const int v = tt == 0 ? a + c : tt == 1 ? a - c : tt == 2 ? b + d : b - d;
                        in[cc] = (int)(int16_t)v;
                    }
                    in[4] = in[5] = in[6] = in[7] = 0;
                    idct8<false>(in, out);
                    *(int4*)(tmp + 128 + tt * 32 + rr * 8) = make_int4(out[0], out[1], out[2], out[3]);
                    *(int4*)(tmp + 128 + tt * 32 + rr * 8 + 4) = make_int4(out[4], out[5], out[6], out[7]);
                }
            }
            __syncwarp();
            if (active) {   // D: column pass, thread = column
                uint8_t* dst = s_tiles + m * IC_MCU420 + (4 + ch * 4) * 64;
#pragma unroll
                for (int tt = 0; tt < 4; ++tt) {
                    int in[8], out[8];
#pragma unroll
                    for (int r = 0; r < 4; ++r) in[r] = tmp[128 + tt * 32 + r * 8 + t];
                    in[4] = in[5] = in[6] = in[7] = 0;
                    idct8<true>(in, out);
#pragma unroll
                    for (int r = 0; r < 8; ++r) dst[tt * 64 + r * 8 + t] = (uint8_t)out[r];
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- stage 2: colour conversion + channel adaptation (H*Convert :2528-2823, :3763-3801)
    const int W = im.width, H = im.height, rc = im.req_comps;
    if (st == YH2V2) {
        const int row = tid >> 4, m = tid & 15;
        const int y = mrow * 16 + row, x = (g0 + m) * 16;
        if (m < nm && y < H && x < W) {
            const uint8_t* tb = s_tiles + m * IC_MCU420 + (row >> 3) * 128 + (row & 7) * 8;
            if (rc == 3) colour420<3>(tb, im.out + ((size_t)y * W + x) * 3, min(16, W - x));
            else if (rc == 4) colour420<4>(tb, im.out + ((size_t)y * W + x) * 4, min(16, W - x));
            else colour420<1>(tb, im.out + ((size_t)y * W + x), min(16, W - x));
        }
    } else {
        const int mw = st == YH2V1 ? 16 : 8, mh = st == YH1V2 ? 16 : 8;
        const int rw = nm * mw;
        for (int p = tid; p < rw * mh; p += IC_THREADS) {
            const int row = p / rw, xx = p - row * rw;
            const int y = mrow * mh + row, x = g0 * mw + xx;
            if (y >= H || x >= W) continue;
            int Y, cb = 128, cr = 128;
            const uint8_t* tb;
            switch (st) {
            case GRAYSCALE: tb = s_tiles + (xx >> 3) * 64; Y = tb[row * 8 + (xx & 7)]; break;
            case YH1V1: { tb = s_tiles + (xx >> 3) * 192; const int o = row * 8 + (xx & 7); Y = tb[o]; cb = tb[64 + o]; cr = tb[128 + o]; break; }
            case YH2V1: { tb = s_tiles + (xx >> 4) * 256; const int lx = xx & 15;
                Y = tb[(lx >> 3) * 64 + row * 8 + (lx & 7)]; cb = tb[128 + row * 8 + (lx >> 1)]; cr = tb[192 + row * 8 + (lx >> 1)]; break; }
            default: { tb = s_tiles + (xx >> 3) * 256; const int lx = xx & 7;      // YH1V2
                Y = tb[(row >> 3) * 64 + (row & 7) * 8 + lx]; cb = tb[128 + (row >> 1) * 8 + lx]; cr = tb[192 + (row >> 1) * 8 + lx]; break; }
            }
            uint8_t* d = im.out + ((size_t)y * W + x) * rc;
            if (im.comps == 1) {
                if (rc == 1) d[0] = (uint8_t)Y;
                else { d[0] = d[1] = d[2] = (uint8_t)Y; if (rc == 4) d[3] = 255; }
            } else {
                int r, g, b; ycc_to_rgb(Y, cb, cr, r, g, b);
                if (rc == 1) d[0] = (uint8_t)((r * 19595 + g * 38470 + b * 7471 + 32768) >> 16);
                else { d[0] = (uint8_t)r; d[1] = (uint8_t)g; d[2] = (uint8_t)b; if (rc == 4) d[3] = 255; }
            }
        }
    }
}

// ---- host: marker layer ----------------------------------------------------------------------
struct ByteSrc {    // get_char semantics: past the end yields FF D9 FF D9 ... (jpegload.d:640-655)
    const uint8_t* p; size_t len, pos; int tem;
    uint32_t next() { if (pos >= len) { int t = tem; tem ^= 1; return t ? 0xD9 : 0xFF; } return p[pos++]; }
    uint32_t u16() { uint32_t a = next(); return (a << 8) | next(); }
};

struct HostHuff { bool valid = false; uint8_t num[17]; uint8_t val[256]; };

struct HostScan {                // one scan of a progressive file (init_progressive, jpegload.d:3609-3671)
    int ncomp = 0, comp[4] = {0}, dc_tab[4] = {0}, ac_tab[4] = {0};      // tables: indices into `huff` below
    int ss = 0, se = 0, ah = 0, al = 0, restart_interval = 0;
    int mcus_per_row = 0, mcus_per_col = 0;
    size_t data_start = 0, data_end = 0;
    HostHuff huff[8];            // the tables as they stood at this SOS (DHT may redefine them between scans)
};

struct Parsed {
    bool ok = false;
    int unsupported = 0;         // 2 non-interleaved multi-scan sequential: a valid JPEG that is not on this path
    bool progressive = false;
    std::vector<HostScan> scans;
    int width = 0, height = 0, comps = 0;
    int h_samp[4] = {0}, v_samp[4] = {0}, quant_sel[4] = {0}, ident[4] = {0};
    HostHuff huff[8];
    bool quant_valid[4] = {false, false, false, false};
    int16_t quant[4][64];
    int restart_interval = 0;
    int comps_in_scan = 0, comp_list[4] = {0}, dc_tab[4] = {0}, ac_tab[4] = {0};
    size_t scan_start = 0;
    float ppiX, ppiY, par;
    int scan_type = 0, mcus_per_row = 0, mcus_per_col = 0, blocks_per_mcu = 0, tiles_per_mcu = 0, mcu_org[10];
};

float inches_to_meters(float x) { return x / 39.37007874f; }      // convertInchesToMeters, types.d:127 (a division there too)
uint16_t rd16(const uint8_t* p, bool le) { return le ? (uint16_t)(p[0] | (p[1] << 8)) : (uint16_t)((p[0] << 8) | p[1]); }
uint32_t rd32(const uint8_t* p, bool le) { return le ? ((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24))
                                                      : (((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]); }

// Walks markers until SOFn/SOI/EOI/SOS (process_markers, jpegload.d:1578-1845). Returns marker or -1.
int walk_markers(ByteSrc& s, Parsed& P)
{
    for (;;) {
        uint32_t c;
        do {
            do { c = s.next(); } while (c != 0xFF);
            do { c = s.next(); } while (c == 0xFF);
        } while (c == 0);
        switch (c) {
        case 0xC0: case 0xC1: case 0xC2: case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB:
        case 0xCD: case 0xCE: case 0xCF: case 0xD8: case 0xD9: case 0xDA:
            return (int)c;
        case 0xC4: {  // DHT (:1177-1268)
            uint32_t left = s.u16();
            if (left < 2) return -1;
            left -= 2;
            while (left) {
                int index = (int)s.next();
                uint8_t num[17]; num[0] = 0; int count = 0;
                for (int i = 1; i <= 16; ++i) { num[i] = (uint8_t)s.next(); count += num[i]; }
                if (count > 255) return -1;
                uint8_t val[256]; memset(val, 0, 256);
                for (int i = 0; i < count; ++i) val[i] = (uint8_t)s.next();
                uint32_t used = 1 + 16 + (uint32_t)count;
                if (left < used) return -1;
                left -= used;
                index = (index & 0x0F) + ((index & 0x10) >> 4) * 4;
                if (index >= 8) return -1;
                P.huff[index].valid = true;
                memcpy(P.huff[index].num, num, 17); memcpy(P.huff[index].val, val, 256);
            }
            break; }
        case 0xCC: return -1;   // arithmetic coding
        case 0xDB: {  // DQT (:1272-1340)
            uint32_t left = s.u16();
            if (left < 2) return -1;
            left -= 2;
            while (left) {
                int n = (int)s.next(); int prec = n >> 4; n &= 15;
                if (n >= 4) return -1;
                P.quant_valid[n] = true;
                for (int i = 0; i < 64; ++i) { uint32_t t = s.next(); if (prec) t = (t << 8) + s.next(); P.quant[n][i] = (int16_t)t; }
                uint32_t used = 65 + (prec ? 64 : 0);
                if (left < used) return -1;
                left -= used;
            }
            break; }
        case 0xDD:   // DRI (:1445-1463)
            if (s.u16() != 4) return -1;
            P.restart_interval = (int)s.u16();
            break;
        case 0xE0: {  // APP0 / JFIF density (:1642-1702)
            uint32_t left = s.u16();
            if (left < 7) return -1;
            left -= 2;
            uint8_t id[5]; for (auto& b : id) b = (uint8_t)s.next();
            left -= 5;
            static const uint8_t JFIF[5] = {0x4A, 0x46, 0x49, 0x46, 0x00};
            if (!memcmp(id, JFIF, 5) && left >= 7) {
                s.u16();
                uint32_t units = s.next(); int Xd = (int)s.u16(), Yd = (int)s.u16();
                left -= 7;
                P.par = (float)(Xd / (double)Yd);
                if (units == 0) { P.ppiX = -1; P.ppiY = -1; }
                else if (units == 1) { P.ppiX = (float)Xd; P.ppiY = (float)Yd; }
                else if (units == 2) { P.ppiX = inches_to_meters(Xd * 100.0f); P.ppiY = inches_to_meters(Yd * 100.0f); }
            }
            while (left) { s.next(); --left; }
            break; }
        case 0xE1: {  // APP1 / EXIF resolution (:1703-1800), bounds-checked
            uint32_t left = s.u16();
            if (left < 2) return -1;
            left -= 2;
            std::vector<uint8_t> ex(left ? left : 1);
            for (uint32_t i = 0; i < left; ++i) ex[i] = (uint8_t)s.next();
            static const uint8_t EXIF[6] = {0x45, 0x78, 0x69, 0x66, 0, 0};
            if (left >= 14 && !memcmp(ex.data(), EXIF, 6)) {
                const uint8_t* tiff = ex.data() + 6; const uint32_t tlen = left - 6;
                uint16_t bo = rd16(tiff, false);
                if (bo != 0x4949 && bo != 0x4D4D) return -1;
                bool le = bo == 0x4949;
                if (rd16(tiff + 2, le) != 42) return -1;
                uint32_t off = rd32(tiff + 4, le);
                double rx = 72, ry = 72; int unit = 2;
                int ifds = 0;
                while (off != 0) {
                    if (++ifds > 64) return -1;          // a next-IFD offset that points back at itself would spin forever (the reference does)
                    if (off > left || (uint64_t)off + 2 > tlen) return -1;
                    const uint8_t* q = tiff + off;
                    uint32_t ne = rd16(q, le); q += 2;
                    if ((uint64_t)(q - tiff) + (uint64_t)ne * 12 + 4 > tlen) return -1;
                    for (uint32_t e = 0; e < ne; ++e, q += 12) {
                        uint32_t tag = rd16(q, le), vo = rd32(q + 8, le);
                        if (tag == 282 || tag == 283) {
                            if ((uint64_t)vo + 8 > tlen) return -1;
                            double num = rd32(tiff + vo, le), den = rd32(tiff + vo + 4, le);
                            if (tag == 282) rx = num / den; else ry = num / den;
                        }
                        if (tag == 296) unit = (int)vo;
                    }
                    off = rd32(q, le);
                }
                if (unit == 2) { P.ppiX = (float)rx; P.ppiY = (float)ry; P.par = (float)(rx / ry); }
                else if (unit == 3) { P.ppiX = inches_to_meters((float)(rx * 100)); P.ppiY = inches_to_meters((float)(ry * 100)); P.par = (float)(rx / ry); }
            }
            break; }
        case 0xD0: case 0xD1: case 0xD2: case 0xD3: case 0xD4: case 0xD5: case 0xD6: case 0xD7: return -1;
        case 0xC8: case 0x01: return -1;
        default: {   // skip_variable_marker (:1408-1432)
            uint32_t left = s.u16();
            if (left < 2) return -1;
            left -= 2;
            while (left) { s.next(); --left; }
            break; }
        }
    }
}

bool parse_jpeg(const uint8_t* data, size_t len, Parsed& P)
{
    P.ppiX = P.ppiY = P.par = std::nanf("");      // D float.init: never assigned without JFIF/EXIF
    ByteSrc s{data, len, 0, 0};
    // the reference pre-fetches 4 bytes into its bit buffer before anything else (initit :2066-2071);
    // with the stream shorter than that the padding toggles -- reproduce the toggle count
    // locate_soi_marker (:1851-1895)
    uint32_t last = s.next(), cur = s.next();
    if (!(last == 0xFF && cur == 0xD8)) {
        uint32_t left = 4096;
        for (;;) {
            if (--left == 0) return false;
            last = cur; cur = s.next();
            if (last == 0xFF) { if (cur == 0xD8) break; if (cur == 0xD9) return false; }
        }
        size_t save = s.pos; int st = s.tem;
        uint32_t nx = s.next(); s.pos = save; s.tem = st;
        if (nx != 0xFF) return false;
    }
    int c = walk_markers(s, P);
    if (c == 0xC2) P.progressive = true;            // locate_sof_marker (:1921-1924)
    else if (c != 0xC0 && c != 0xC1) return false;
    // read_sof_marker (:1343-1405)
    uint32_t left = s.u16();
    if (s.next() != 8) return false;
    P.height = (int)s.u16(); if (P.height < 1 || P.height > 16384) return false;
    P.width = (int)s.u16(); if (P.width < 1 || P.width > 16384) return false;
    P.comps = (int)s.next(); if (P.comps > 4) return false;
    if (left != (uint32_t)(P.comps * 3 + 8)) return false;
    for (int i = 0; i < P.comps; ++i) { P.ident[i] = (int)s.next(); uint32_t hv = s.next(); P.h_samp[i] = hv >> 4; P.v_samp[i] = hv & 15; P.quant_sel[i] = (int)s.next(); }
    // init_frame (:3130-3268)
    if (P.comps == 1) { if (P.h_samp[0] != 1 || P.v_samp[0] != 1) return false; P.scan_type = GRAYSCALE; }
    else if (P.comps == 3) {
        if (P.h_samp[1] != 1 || P.v_samp[1] != 1 || P.h_samp[2] != 1 || P.v_samp[2] != 1) return false;
        if (P.h_samp[0] == 1 && P.v_samp[0] == 1) P.scan_type = YH1V1;
        else if (P.h_samp[0] == 2 && P.v_samp[0] == 1) P.scan_type = YH2V1;
        else if (P.h_samp[0] == 1 && P.v_samp[0] == 2) P.scan_type = YH1V2;
        else if (P.h_samp[0] == 2 && P.v_samp[0] == 2) P.scan_type = YH2V2;
        else return false;
    } else return false;
    {
        static const int bpm[5] = {1, 3, 4, 4, 6};
        if ((P.width + (P.scan_type == YH2V1 || P.scan_type == YH2V2 ? 15 : 7)) / (P.scan_type == YH2V1 || P.scan_type == YH2V2 ? 16 : 8) * bpm[P.scan_type] > 8192) return false;
    }
    if (P.progressive) {
        // init_progressive (:3587-3684): every scan up to EOI; the geometry of the output stage is that of one
        // interleaved scan over all components (:3675-3681)
        const int max_h = P.h_samp[0], max_v = P.v_samp[0];
        int h_blocks[4], v_blocks[4];
        for (int ci = 0; ci < P.comps; ++ci) {
            h_blocks[ci] = (((P.width * P.h_samp[ci]) + (max_h - 1)) / max_h + 7) / 8;
            v_blocks[ci] = (((P.height * P.v_samp[ci]) + (max_v - 1)) / max_v + 7) / 8;
        }
        const int full_mpr = (((P.width + 7) / 8) + (max_h - 1)) / max_h, full_mpc = (((P.height + 7) / 8) + (max_v - 1)) / max_v;
        for (;;) {
            c = walk_markers(s, P);
            if (c == 0xD9) break;
            if (c != 0xDA) return false;
            HostScan S;
            uint32_t l2 = s.u16();
            const int n = (int)s.next();
            S.ncomp = n;
            l2 -= 3;
            if (l2 != (uint32_t)(n * 2 + 3) || n < 1 || n > 4) return false;
            for (int i = 0; i < n; ++i) {
                const int cc = (int)s.next(), tc = (int)s.next();
                int ci = 0;
                for (; ci < P.comps; ++ci) if (cc == P.ident[ci]) break;
                if (ci >= P.comps) return false;
                S.comp[i] = ci; S.dc_tab[i] = (tc >> 4) & 15; S.ac_tab[i] = (tc & 15) + 4;
                l2 -= 2;
            }
            S.ss = (int)s.next(); S.se = (int)s.next();
            { const uint32_t a = s.next(); S.ah = (int)(a >> 4); S.al = (int)(a & 15); }
            l2 -= 3;
            while (l2) { s.next(); --l2; }
            if (s.pos > len) return false;
            // calc_mcu_block_order for this scan (:3038-3090)
            if (n == 1) { S.mcus_per_row = h_blocks[S.comp[0]]; S.mcus_per_col = v_blocks[S.comp[0]]; }
            else {
                S.mcus_per_row = full_mpr; S.mcus_per_col = full_mpc;
                int nb = 0;
                for (int i = 0; i < n; ++i) nb += P.h_samp[S.comp[i]] * P.v_samp[S.comp[i]];
                if (nb > 10) return false;
            }
            // check_huff_tables / check_quant_tables (:2990-3035)
            for (int i = 0; i < n; ++i) {
                if (S.ss == 0 && (S.dc_tab[i] >= 8 || !P.huff[S.dc_tab[i]].valid)) return false;
                if (S.se > 0 && (S.ac_tab[i] >= 8 || !P.huff[S.ac_tab[i]].valid)) return false;
                if (P.quant_sel[S.comp[i]] >= 4 || !P.quant_valid[P.quant_sel[S.comp[i]]]) return false;
            }
            // the scan's own checks (:3620-3645)
            if (S.ss > S.se || S.se > 63) return false;
            if (S.ss == 0) { if (S.se) return false; }
            else if (n != 1) return false;
            if (S.ah != 0 && S.al != S.ah - 1) return false;
            if (S.al > 13) return false;            // shifts of a 16-bit coefficient (T.81 allows 0..13)
            for (int t = 0; t < 8; ++t) S.huff[t] = P.huff[t];
            S.restart_interval = P.restart_interval;
            S.data_start = s.pos;
            // the entropy data runs up to the next marker that is neither a stuffed FF00 nor RSTn
            size_t e = s.pos;
            while (e + 1 < len && !(data[e] == 0xFF && data[e + 1] != 0x00 && !(data[e + 1] >= 0xD0 && data[e + 1] <= 0xD7))) ++e;
            if (e + 1 >= len) e = len;
            S.data_end = e;
            s.pos = e;
            P.scans.push_back(S);
            if (P.scans.size() > 1024) return false;
        }
        // the output stage: one interleaved MCU grid over all components
        P.comps_in_scan = P.comps;
        for (int i = 0; i < P.comps; ++i) P.comp_list[i] = i;
        if (P.comps == 1) { P.mcus_per_row = h_blocks[0]; P.mcus_per_col = v_blocks[0]; P.blocks_per_mcu = 1; P.mcu_org[0] = 0; }
        else {
            P.mcus_per_row = full_mpr; P.mcus_per_col = full_mpc;
            P.blocks_per_mcu = 0;
            for (int ci = 0; ci < P.comps; ++ci) { int nb = P.h_samp[ci] * P.v_samp[ci]; while (nb--) { if (P.blocks_per_mcu >= 10) return false; P.mcu_org[P.blocks_per_mcu++] = ci; } }
        }
        P.tiles_per_mcu = P.scan_type == YH2V2 ? 12 : P.blocks_per_mcu;
        for (int ci = 0; ci < P.comps; ++ci) if (P.quant_sel[ci] >= 4 || !P.quant_valid[P.quant_sel[ci]]) return false;
        P.ok = true;
        return true;
    }
    // init_scan (:3093-3127): next SOS
    c = walk_markers(s, P);
    if (c != 0xDA) return false;
    // read_sos_marker (:1466-1543)
    left = s.u16();
    int n = (int)s.next();
    P.comps_in_scan = n;
    left -= 3;
    if (left != (uint32_t)(n * 2 + 3) || n < 1 || n > 4) return false;
    for (int i = 0; i < n; ++i) {
        int cc = (int)s.next(), tc = (int)s.next();
        left -= 2;
        int ci = 0;
        for (; ci < P.comps; ++ci) if (cc == P.ident[ci]) break;
        if (ci >= P.comps) return false;
        P.comp_list[i] = ci; P.dc_tab[ci] = (tc >> 4) & 15; P.ac_tab[ci] = (tc & 15) + 4;
    }
    s.next(); s.next(); s.next();
    left -= 3;
    while (left) { s.next(); --left; }
    if (s.pos > len) return false;
    if (P.comps_in_scan != P.comps) { P.unsupported = 2; return false; }  // non-interleaved multi-scan baseline: unsupported
    // calc_mcu_block_order (:3038-3090)
    int max_h = P.h_samp[0], max_v = P.v_samp[0];
    if (P.comps == 1) { P.mcus_per_row = (P.width + 7) / 8; P.mcus_per_col = (P.height + 7) / 8; P.blocks_per_mcu = 1; P.mcu_org[0] = P.comp_list[0]; }
    else {
        P.mcus_per_row = (((P.width + 7) / 8) + (max_h - 1)) / max_h;
        P.mcus_per_col = (((P.height + 7) / 8) + (max_v - 1)) / max_v;
        P.blocks_per_mcu = 0;
        for (int k = 0; k < n; ++k) { int ci = P.comp_list[k]; int nb = P.h_samp[ci] * P.v_samp[ci]; while (nb--) { if (P.blocks_per_mcu >= 10) return false; P.mcu_org[P.blocks_per_mcu++] = ci; } }
    }
    P.tiles_per_mcu = P.scan_type == YH2V2 ? 12 : P.blocks_per_mcu;
    for (int i = 0; i < n; ++i) {
        int ci = P.comp_list[i];
        if (P.dc_tab[ci] >= 8 || !P.huff[P.dc_tab[ci]].valid) return false;
        if (P.ac_tab[ci] >= 8 || !P.huff[P.ac_tab[ci]].valid) return false;
        if (P.quant_sel[ci] >= 4 || !P.quant_valid[P.quant_sel[ci]]) return false;
    }
    P.scan_start = s.pos;
    P.ok = true;
    return true;
}

void build_table(const HostHuff& h, HuffTable& T, bool is_dc)
{
    memset(&T, 0, sizeof(T));
    memcpy(T.val, h.val, 256);
    T.is_dc = is_dc ? 1 : 0;
    // what is not a code word reads as one bit that moves a DC position on and ends the block from an AC position
    const uint16_t bad_e = (uint16_t)((is_dc ? 1 : HT_KINC_EOB) << 9);
    for (auto& v : T.l1) v = bad_e;
    for (auto& sub : T.l2) for (auto& v : sub) v = bad_e;
    int code = 0, k = 0, nsubs = 0;
    int sub_of[1 << HT_L1_BITS];
    for (int& v : sub_of) v = -1;
    for (int l = 1; l <= 16; ++l) {
        T.valptr[l] = k; T.mincode[l] = code;
        for (int i = 0; i < h.num[l]; ++i, ++k, ++code) {
            if (code >= (1 << l)) continue;          // over-subscribed table: the code does not exist
            const int sym = h.val[k];
            if (l <= HUFF_FAST) {
                int shift = HUFF_FAST - l;
                for (int f = 0; f < (1 << shift); ++f) T.fast[(code << shift) | f] = (uint16_t)((l << 8) | sym);
            }
            // a DC symbol is a size 0..15; anything else is not a code word for the decoders
            const bool valid = !is_dc || sym <= 15;
            const int size = sym & 15, run = sym >> 4;
            const int kinc = is_dc ? 1 : (size ? run + 1 : (run == 15 ? 16 : HT_KINC_EOB));
            const uint16_t e = valid ? (uint16_t)(l | (size << 5) | (kinc << 9)) : bad_e;
            if (l <= HT_L1_BITS) {
                const int shift = HT_L1_BITS - l;
                for (int f = 0; f < (1 << shift); ++f) T.l1[(code << shift) | f] = e;
            } else {
                const int prefix = code >> (l - HT_L1_BITS);
                if (sub_of[prefix] < 0) sub_of[prefix] = nsubs < HT_SUBS ? nsubs++ : HT_SUBS;
                if (sub_of[prefix] == HT_SUBS) { T.l1[prefix] = (uint16_t)(HT_LONG | 7); continue; }
                T.l1[prefix] = (uint16_t)(HT_LONG | sub_of[prefix]);
                const int rest = code & ((1 << (l - HT_L1_BITS)) - 1), shift = HT_L1_BITS + HT_L2_BITS - l;
                for (int f = 0; f < (1 << shift); ++f) T.l2[sub_of[prefix]][(rest << shift) | f] = e;
            }
        }
        T.maxcode[l] = h.num[l] ? code - 1 : -1;
        code <<= 1;
    }
}

inline size_t al(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
inline double now_ms() { using namespace std::chrono; return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count(); }

} // namespace

namespace gb {

gb200_batch* jpeg_decode_batch(int n, const uint8_t* const* files, const size_t* lens,
                               const uint8_t* const* files_dev, int req_comps_in, cudaStream_t st)
{
    if (!ensure_device()) return nullptr;
    if (n < 0) { set_error("jpeg_decode_batch: negative count"); return nullptr; }
    gb200_batch* B = new gb200_batch;
    B->stream = st;
    B->images.resize((size_t)n);
    for (auto& D : B->images) { memset(&D, 0, sizeof(D)); D.ppmX = D.ppmY = D.pixelAspectRatio = -1; }
    if (req_comps_in != -1 && req_comps_in != 1 && req_comps_in != 3 && req_comps_in != 4) return B;   // every image fails (:3727)
    double t0 = now_ms();
    std::vector<Parsed> P((size_t)n);
    std::map<std::string, int> table_ids;
    std::vector<HuffTable> tables;
    std::vector<int> live;
    size_t out_total = 0;
    std::vector<size_t> out_off((size_t)n, 0);
    for (int i = 0; i < n; ++i) {
        if (!files[i] || !parse_jpeg(files[i], lens[i], P[i])) continue;
        live.push_back(i);
        int rc = req_comps_in < 0 ? P[i].comps : req_comps_in;
        out_off[i] = out_total;
        out_total += al((size_t)P[i].width * P[i].height * rc);
    }
    B->host_parse_ms = now_ms() - t0;
    uint8_t* d_out = nullptr;
    if (out_total) { d_out = (uint8_t*)dev_alloc(out_total); if (!d_out) { delete B; return nullptr; } B->device_allocs.push_back(d_out); }

    // process in chunks that bound the coefficient scratch
    const size_t SCRATCH_BUDGET = (size_t)24 << 30;
    size_t li = 0;
    std::vector<int> final_ok((size_t)n, 0);
    while (li < live.size()) {
        size_t scratch = 0, lj = li;
        std::vector<size_t> coef_off, zag_off, file_off;
        size_t file_total = 0, zag_total = 0;
        std::vector<size_t> plane_off;            // progressive images: whole-image coefficient planes after the MCU-ordered ones
        size_t plane_total = 0;
        auto plane_bytes = [](const Parsed& p) {
            if (!p.progressive) return (size_t)0;
            const int mw = p.scan_type == YH2V1 || p.scan_type == YH2V2 ? 16 : 8, mh = p.scan_type == YH1V2 || p.scan_type == YH2V2 ? 16 : 8;
            const size_t mr = (size_t)(p.width + mw - 1) / mw, mc = (size_t)(p.height + mh - 1) / mh;   // m_max_mcus_per_row / col
            size_t t = 0;
            for (int c = 0; c < p.comps; ++c) t += al(mr * p.h_samp[c] * mc * p.v_samp[c] * 128);
            return t;
        };
        while (lj < live.size()) {
            const Parsed& p = P[live[lj]];
            size_t mcus = (size_t)p.mcus_per_row * p.mcus_per_col;
            size_t need = al(mcus * p.blocks_per_mcu * 128);
            const size_t pneed = plane_bytes(p);
            if (lj > li && scratch + plane_total + need + pneed > SCRATCH_BUDGET) break;
            coef_off.push_back(scratch);
            plane_off.push_back(plane_total); plane_total += pneed;
            scratch += need;
            zag_off.push_back(zag_total); zag_total += al(mcus * p.blocks_per_mcu);
            file_off.push_back(file_total); file_total += al(lens[live[lj]] + 16);
            ++lj;
        }
        const int m = (int)(lj - li);
        DevBuf d_scratch(scratch), d_zag(zag_total + 256), d_files(files_dev ? 256 : file_total), d_status(sizeof(int) * (size_t)m),
               d_planes(plane_total + 256);
        if (!d_scratch.p || !d_zag.p || !d_files.p || !d_status.p || !d_planes.p) { delete B; return nullptr; }
        std::vector<ProgImage> pimgs; std::vector<ProgScan> pscans; std::vector<Segment> psegs;
        std::vector<uint32_t> pblk_base(1, 0);
        std::vector<JpegImage> imgs((size_t)m);
        std::vector<Segment> segs;
        std::vector<uint32_t> cta_base((size_t)m + 1, 0);     // fused IDCT+colour kernel: CTAs per image (prefix)
        uint8_t* h_stage = nullptr;
        if (!files_dev) { h_stage = (uint8_t*)pinned_alloc(file_total); if (!h_stage) { delete B; return nullptr; } }
        std::vector<int> host_fail((size_t)m, 0);
        std::vector<HostCopy> hcopies;
        for (int k = 0; k < m; ++k) {
            const int i = live[li + k];
            const Parsed& p = P[i];
            JpegImage& J = imgs[k];
            memset(&J, 0, sizeof(J));
            if (files_dev) J.data = files_dev[i];
            else { hcopies.push_back(HostCopy{h_stage + file_off[k], files[i], lens[i]}); J.data = d_files.as<uint8_t>() + file_off[k]; }
            J.data_len = (uint32_t)lens[i];
            J.width = p.width; J.height = p.height; J.scan_type = p.scan_type; J.comps = p.comps;
            J.mcus_per_row = p.mcus_per_row; J.mcus_per_col = p.mcus_per_col; J.blocks_per_mcu = p.blocks_per_mcu; J.tiles_per_mcu = p.tiles_per_mcu;
            for (int b = 0; b < p.blocks_per_mcu; ++b) J.mcu_org[b] = p.mcu_org[b];
            auto table_id = [&](const HostHuff& hh, bool is_ac) {
                std::string key((const char*)hh.num, 17); key.append((const char*)hh.val, 256); key.push_back(is_ac ? 'A' : 'D');
                auto it = table_ids.find(key);
                if (it != table_ids.end()) return it->second;
                const int id = (int)tables.size();
                tables.emplace_back(); build_table(hh, tables.back(), !is_ac); table_ids[key] = id;
                return id;
            };
            for (int c = 0; c < p.comps; ++c) {
                memcpy(J.quant[c], p.quant[p.quant_sel[c]], 128);
                if (p.progressive) continue;                  // tables belong to the scans
                for (int which = 0; which < 2; ++which)
                    (which ? J.ac_tab[c] : J.dc_tab[c]) = table_id(p.huff[which ? p.ac_tab[c] : p.dc_tab[c]], which != 0);
            }
            J.coefs = (int16_t*)(d_scratch.as<uint8_t>() + coef_off[k]);
            J.blk_zag = d_zag.as<uint8_t>() + zag_off[k];
            J.samples = nullptr;
            J.out = d_out + out_off[i];
            J.req_comps = req_comps_in < 0 ? p.comps : req_comps_in;
            J.restart_interval = p.restart_interval;
            const int total_mcus = p.mcus_per_row * p.mcus_per_col;
            {
                const int NM = p.scan_type == YH2V2 ? 16 : 32;
                cta_base[k + 1] = cta_base[k] + (uint32_t)((p.mcus_per_row + NM - 1) / NM) * (uint32_t)p.mcus_per_col;
            }
            const uint8_t* f = files[i]; const size_t flen = lens[i];
            if (p.progressive) {
                // whole-image coefficient planes (coeff_buf_open, :3601-3604) and the scans that fill them
                ProgImage PI; memset(&PI, 0, sizeof(PI));
                PI.image = k; PI.first_scan = (int)pscans.size(); PI.nscans = (int)p.scans.size();
                const int mw = p.scan_type == YH2V1 || p.scan_type == YH2V2 ? 16 : 8, mh = p.scan_type == YH1V2 || p.scan_type == YH2V2 ? 16 : 8;
                const int mr = (p.width + mw - 1) / mw, mc = (p.height + mh - 1) / mh;
                size_t off = plane_off[k];
                for (int c = 0; c < p.comps; ++c) {
                    PI.h[c] = p.h_samp[c]; PI.v[c] = p.v_samp[c];
                    PI.num_x[c] = mr * p.h_samp[c]; PI.num_y[c] = mc * p.v_samp[c];
                    PI.plane[c] = (int16_t*)(d_planes.as<uint8_t>() + off);
                    off += al((size_t)PI.num_x[c] * PI.num_y[c] * 128);
                }
                for (const HostScan& hs : p.scans) {
                    ProgScan S; memset(&S, 0, sizeof(S));
                    S.ncomp = hs.ncomp; S.ss = hs.ss; S.se = hs.se; S.ah = hs.ah; S.al = hs.al; S.mcus_per_row = hs.mcus_per_row;
                    for (int q = 0; q < hs.ncomp && q < 3; ++q) {
                        S.comp[q] = hs.comp[q];
                        S.dc_tab[q] = hs.ss == 0 ? table_id(hs.huff[hs.dc_tab[q]], false) : 0;
                        S.ac_tab[q] = hs.se > 0 ? table_id(hs.huff[hs.ac_tab[q]], true) : 0;
                    }
                    S.first_seg = (int)psegs.size();
                    // one segment per restart interval (process_restart, :2335-2402), the whole scan without DRI
                    const int scan_mcus = hs.mcus_per_row * hs.mcus_per_col;
                    if (!hs.restart_interval) psegs.push_back(Segment{(int)pimgs.size(), (uint32_t)hs.data_start, (uint32_t)hs.data_end, 0, scan_mcus});
                    else {
                        size_t pos = hs.data_start; int mcu = 0, expect = 0; bool bad = false;
                        while (mcu < scan_mcus) {
                            const int cnt = std::min(hs.restart_interval, scan_mcus - mcu);
                            size_t e = pos;
                            while (e + 1 < hs.data_end && !(f[e] == 0xFF && f[e + 1] != 0x00)) ++e;
                            if (e + 1 >= hs.data_end) e = hs.data_end;
                            psegs.push_back(Segment{(int)pimgs.size(), (uint32_t)pos, (uint32_t)e, mcu, cnt});
                            mcu += cnt;
                            if (mcu >= scan_mcus) break;
                            size_t q = e;
                            while (q < hs.data_end && f[q] == 0xFF) ++q;
                            if (q >= hs.data_end || q == e || f[q] != (uint8_t)(0xD0 + expect)) { bad = true; break; }
                            expect = (expect + 1) & 7;
                            pos = q + 1;
                        }
                        if (bad) host_fail[k] = 1;
                    }
                    S.nsegs = (int)psegs.size() - S.first_seg;
                    pscans.push_back(S);
                }
                pimgs.push_back(PI);
                pblk_base.push_back(pblk_base.back() + (uint32_t)total_mcus * (uint32_t)p.blocks_per_mcu);
                continue;
            }
            // segments: the whole scan, or one per restart interval (process_restart, :2335-2402)
            if (!p.restart_interval) segs.push_back(Segment{k, (uint32_t)p.scan_start, (uint32_t)flen, 0, total_mcus});
            else {
                size_t pos = p.scan_start; int mcu = 0, expect = 0; bool bad = false;
                while (mcu < total_mcus) {
                    int cnt = std::min(p.restart_interval, total_mcus - mcu);
                    // find the end of this interval's data: the next marker that is not FF00
                    size_t e = pos;
                    while (e + 1 < flen && !(f[e] == 0xFF && f[e + 1] != 0x00)) ++e;
                    if (e + 1 >= flen) e = flen;
                    segs.push_back(Segment{k, (uint32_t)pos, (uint32_t)e, mcu, cnt});
                    mcu += cnt;
                    if (mcu >= total_mcus) break;
                    // the next interval must start with RST(expect), possibly preceded by fill FFs
                    size_t q = e;
                    while (q < flen && f[q] == 0xFF) ++q;
                    if (q >= flen || q == e || f[q] != (uint8_t)(0xD0 + expect)) { bad = true; break; }
                    expect = (expect + 1) & 7;
                    pos = q + 1;
                }
                if (bad) host_fail[k] = 1;
            }
        }
        // entropy segments long enough to be worth cutting into chunks go to the self-synchronising decoder; the
        // others are decoded by one thread each into zero-filled blocks
        std::vector<LongSeg> longsegs;
        std::vector<char> needs_zero((size_t)m, 0);
        uint32_t total_chunks = 0, total_sync_ctas = 0, total_tiles = 0; size_t clean_total = 0;
        {
            std::vector<Segment> shortsegs;
            for (const Segment& sg : segs) {
                const uint32_t len = sg.end - sg.start;
                if (len < JS_LONG_MIN || host_fail[sg.image]) { shortsegs.push_back(sg); needs_zero[sg.image] = 1; continue; }
                LongSeg L;
                L.image = sg.image; L.in_start = sg.start; L.in_end = sg.end; L.first_mcu = sg.first_mcu; L.num_mcus = sg.num_mcus;
                L.chunk_base = total_chunks; L.nchunks = (len + JS_CHUNK_BYTES - 1) / JS_CHUNK_BYTES;
                L.cta_base = total_sync_ctas;
                L.clean_off = clean_total;
                L.tile_base = total_tiles; L.ntiles = (sg.end - (sg.start & ~15u) + JU_TILE - 1) / JU_TILE; total_tiles += L.ntiles;
                total_chunks += (L.nchunks + JW_CTA - 1) / JW_CTA * JW_CTA;        // the write kernel's CTAs never span segments
                total_sync_ctas += (L.nchunks + JS_OWN - 1) / JS_OWN;
                clean_total += al((size_t)len + JS_PAD_BYTES + 16, 16);
                longsegs.push_back(L);
            }
            segs.swap(shortsegs);
        }
        const int nlong = (int)longsegs.size();
        DevBuf d_long(sizeof(LongSeg) * ((size_t)nlong + 1)), d_clean(clean_total + 256), d_clen(4 * ((size_t)nlong + 1)),
               d_recs(sizeof(ChunkRec) * ((size_t)total_chunks + 1)), d_bases(sizeof(ChunkBase) * ((size_t)total_chunks + 1)),
               d_entry(sizeof(ChunkState) * ((size_t)total_sync_ctas + 1)), d_unconv(256),
               d_tiles(sizeof(JuTile) * ((size_t)total_tiles + 1)), d_segend(4 * ((size_t)nlong + 1));
        if (!d_tiles.p || !d_segend.p || !d_long.p || !d_clean.p || !d_clen.p || !d_recs.p || !d_bases.p || !d_entry.p || !d_unconv.p) {
            if (h_stage) pinned_free(h_stage); delete B; return nullptr;
        }
        DevBuf d_imgs(sizeof(JpegImage) * (size_t)m), d_segs(sizeof(Segment) * (segs.size() + 1)),
               d_tables(sizeof(HuffTable) * (tables.size() + 1)), d_base(sizeof(int) * ((size_t)m + 1));
        if (!d_imgs.p || !d_segs.p || !d_tables.p || !d_base.p) { if (h_stage) pinned_free(h_stage); delete B; return nullptr; }
        std::vector<int> st_init((size_t)m);
        for (int k = 0; k < m; ++k) st_init[k] = host_fail[k] ? 0 : 1;
        cudaEvent_t ev[7];
        for (auto& e : ev) cudaEventCreate(&e);
        bool okc = true;
        cudaEventRecord(ev[0], st);
        host_copy_parallel(hcopies.data(), hcopies.size());
        if (!files_dev) okc &= cuda_ok(cudaMemcpyAsync(d_files.p, h_stage, file_total, cudaMemcpyHostToDevice, st), "files", __FILE__, __LINE__);
        okc &= cuda_ok(cudaMemcpyAsync(d_imgs.p, imgs.data(), sizeof(JpegImage) * m, cudaMemcpyHostToDevice, st), "imgs", __FILE__, __LINE__);
        if (!segs.empty()) okc &= cuda_ok(cudaMemcpyAsync(d_segs.p, segs.data(), sizeof(Segment) * segs.size(), cudaMemcpyHostToDevice, st), "segs", __FILE__, __LINE__);
        if (nlong) okc &= cuda_ok(cudaMemcpyAsync(d_long.p, longsegs.data(), sizeof(LongSeg) * nlong, cudaMemcpyHostToDevice, st), "longsegs", __FILE__, __LINE__);
        okc &= cuda_ok(cudaMemcpyAsync(d_tables.p, tables.data(), sizeof(HuffTable) * tables.size(), cudaMemcpyHostToDevice, st), "tables", __FILE__, __LINE__);
        okc &= cuda_ok(cudaMemcpyAsync(d_base.p, cta_base.data(), sizeof(uint32_t) * (m + 1), cudaMemcpyHostToDevice, st), "base", __FILE__, __LINE__);
        okc &= cuda_ok(cudaMemcpyAsync(d_status.p, st_init.data(), sizeof(int) * m, cudaMemcpyHostToDevice, st), "status", __FILE__, __LINE__);
        okc &= dev_fill_async(d_unconv.p, 0, 4, st);
        // images with one-thread segments scatter into zero-filled blocks (the reference zero-fills per block, :2459-2510);
        // the chunk-parallel writer zero-fills its own blocks
        for (int k = 0; k < m; ++k) if (needs_zero[k]) {
            const size_t bytes = (k + 1 < m ? coef_off[k + 1] : scratch) - coef_off[k];
            okc &= dev_fill_async(d_scratch.as<uint8_t>() + coef_off[k], 0, bytes, st);
        }
        // progressive images: scan tables up, planes cleared, one thread per image through all its scans, then the
        // planes are dequantised into the MCU order of the sequential path
        const int nprog = (int)pimgs.size();
        DevBuf d_pimgs(sizeof(ProgImage) * ((size_t)nprog + 1)), d_pscans(sizeof(ProgScan) * (pscans.size() + 1)),
               d_psegs(sizeof(Segment) * (psegs.size() + 1)), d_pblk(sizeof(uint32_t) * ((size_t)nprog + 2));
        if (!d_pimgs.p || !d_pscans.p || !d_psegs.p || !d_pblk.p) okc = false;
        if (nprog && okc) {
            okc &= cuda_ok(cudaMemcpyAsync(d_pimgs.p, pimgs.data(), sizeof(ProgImage) * nprog, cudaMemcpyHostToDevice, st), "prog images", __FILE__, __LINE__);
            if (!pscans.empty()) okc &= cuda_ok(cudaMemcpyAsync(d_pscans.p, pscans.data(), sizeof(ProgScan) * pscans.size(), cudaMemcpyHostToDevice, st), "prog scans", __FILE__, __LINE__);
            if (!psegs.empty()) okc &= cuda_ok(cudaMemcpyAsync(d_psegs.p, psegs.data(), sizeof(Segment) * psegs.size(), cudaMemcpyHostToDevice, st), "prog segments", __FILE__, __LINE__);
            okc &= cuda_ok(cudaMemcpyAsync(d_pblk.p, pblk_base.data(), sizeof(uint32_t) * pblk_base.size(), cudaMemcpyHostToDevice, st), "prog blocks", __FILE__, __LINE__);
        }
        auto run_prog = [&]() {
            if (!nprog || !okc) return;
            okc &= dev_fill_async(d_planes.p, 0, plane_total, st);
            jpeg_prog_kernel<<<(nprog + 31) / 32, 32, 0, st>>>(d_imgs.as<JpegImage>(), d_pimgs.as<ProgImage>(), nprog, d_pscans.as<ProgScan>(),
                                                             d_psegs.as<Segment>(), d_tables.as<HuffTable>(), d_status.as<int>());
            const uint32_t nwarps = pblk_base.back();
            if (nwarps) jpeg_prog_gather_kernel<<<(nwarps + 3) / 4, 128, 0, st>>>(d_imgs.as<JpegImage>(), d_pimgs.as<ProgImage>(), d_pblk.as<uint32_t>(), nprog, d_status.as<int>());
            count_launch(2);
        };
        cudaEventRecord(ev[1], st);
        run_prog();
        const int nsegs = (int)segs.size();
        if (nsegs) {
            jpeg_huffman_kernel<<<(nsegs + 127) / 128, 128, 0, st>>>(d_imgs.as<JpegImage>(), d_segs.as<Segment>(), nsegs, d_tables.as<HuffTable>(), d_status.as<int>());
            count_launch();
        }
        uint32_t h_unconv = 0;
        const JpegImage* dI = d_imgs.as<JpegImage>(); const LongSeg* dL = d_long.as<LongSeg>(); const HuffTable* dT = d_tables.as<HuffTable>();
        uint8_t* clean = d_clean.as<uint8_t>(); uint32_t* clen = d_clen.as<uint32_t>();
        ChunkRec* recs = d_recs.as<ChunkRec>(); ChunkBase* bases = d_bases.as<ChunkBase>(); ChunkState* entry = d_entry.as<ChunkState>();
        uint32_t* unconv = d_unconv.as<uint32_t>();
        const unsigned rg = (total_sync_ctas + 63) / 64;
        auto finish_entropy = [&]() {       // scan + write: everything after the chunk states are final
            jpeg_scan_kernel<<<nlong, 256, 0, st>>>(dI, dL, recs, bases, d_status.as<int>());
            jpeg_write_kernel<<<total_chunks / JW_CTA, JW_CTA, 0, st>>>(dI, dL, nlong, clean, clen, dT, recs, bases, d_status.as<int>());
            count_launch(2);
        };
        if (nlong) {
            // long entropy segments: chunk-parallel self-synchronising decode (jpeg_sync.cuh), no host round trip
            jpeg_unstuff_count_kernel<<<total_tiles, 256, 0, st>>>(dI, dL, nlong, d_tiles.as<JuTile>());
            jpeg_unstuff_scan_kernel<<<nlong, 32, 0, st>>>(dL, d_tiles.as<JuTile>(), d_segend.as<uint32_t>());
            jpeg_unstuff_write_kernel<<<total_tiles, 256, 0, st>>>(dI, dL, nlong, d_tiles.as<JuTile>(), d_segend.as<uint32_t>(), clean, clen);
            count_launch(2);
            cudaEventRecord(ev[4], st);
            jpeg_sync_kernel<<<total_sync_ctas, JS_CTA, 0, st>>>(dI, dL, nlong, clean, clen, dT, recs, entry);
            jpeg_repair_kernel<<<rg, 64, 0, st>>>(dI, dL, nlong, total_sync_ctas, clean, clen, dT, recs, entry, 0, unconv);
            jpeg_repair_kernel<<<rg, 64, 0, st>>>(dI, dL, nlong, total_sync_ctas, clean, clen, dT, recs, entry, 0, unconv);
            jpeg_repair_kernel<<<rg, 64, 0, st>>>(dI, dL, nlong, total_sync_ctas, clean, clen, dT, recs, entry, 1, unconv);
            cudaEventRecord(ev[5], st);
            count_launch(5);
            finish_entropy();
        } else { cudaEventRecord(ev[4], st); cudaEventRecord(ev[5], st); }
        cudaEventRecord(ev[2], st);
        auto run_idct = [&]() {
            uint32_t most = 0;
            for (int k = 0; k < m; ++k) most = std::max(most, cta_base[k + 1] - cta_base[k]);
            for (int k0 = 0; most && k0 < m; k0 += 65535) {          // grid.y is limited to 65535
                const int mk = std::min(65535, m - k0);
                jpeg_idct_colour_kernel<<<dim3(most, (unsigned)mk), IC_THREADS, 0, st>>>(d_imgs.as<JpegImage>() + k0, d_base.as<uint32_t>() + k0, mk, d_status.as<int>() + k0);
                count_launch();
            }
        };
        run_idct();
        cudaEventRecord(ev[3], st);
        // results come back through pinned memory written by a kernel, not through the copy engine (common.h)
        PinnedBuf h_back(sizeof(int) * ((size_t)m + 1));
        if (!h_back.p) okc = false;
        int* const status = h_back.as<int>();
        volatile uint32_t* const h_unconv_p = (volatile uint32_t*)(status + m);
        okc = okc && dev_read_back_async(status, d_status.p, sizeof(int) * m, st);
        okc = okc && dev_read_back_async((void*)h_unconv_p, unconv, 4, st);
        okc &= cuda_ok(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__);
        if (okc) h_unconv = *h_unconv_p;
        // A CTA boundary of the sync kernel was still wrong after two repair rounds (a run of > 248 chunks that never
        // re-synchronises: not seen on real streams): repair until the chain is consistent, then redo the dependent passes.
        for (int round = 0; okc && h_unconv && round < 1 << 16; ++round) {
            okc &= dev_fill_async(unconv, 0, 4, st);
            jpeg_repair_kernel<<<rg, 64, 0, st>>>(dI, dL, nlong, total_sync_ctas, clean, clen, dT, recs, entry, 0, unconv);
            jpeg_repair_kernel<<<rg, 64, 0, st>>>(dI, dL, nlong, total_sync_ctas, clean, clen, dT, recs, entry, 1, unconv);
            count_launch(2);
            okc = okc && dev_read_back_async((void*)h_unconv_p, unconv, 4, st);
            okc &= cuda_ok(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__);
            if (okc) h_unconv = *h_unconv_p;
            if (!h_unconv) {
                okc &= cuda_ok(cudaMemcpyAsync(d_status.p, st_init.data(), sizeof(int) * m, cudaMemcpyHostToDevice, st), "status", __FILE__, __LINE__);
                run_prog();
                if (nsegs) { jpeg_huffman_kernel<<<(nsegs + 127) / 128, 128, 0, st>>>(d_imgs.as<JpegImage>(), d_segs.as<Segment>(), nsegs, d_tables.as<HuffTable>(), d_status.as<int>()); count_launch(); }
                finish_entropy();
                run_idct();
                okc = okc && dev_read_back_async(status, d_status.p, sizeof(int) * m, st);
                okc &= cuda_ok(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__);
            }
        }
        if (h_unconv) okc = false;
        okc &= cuda_ok(cudaGetLastError(), "kernels", __FILE__, __LINE__);
        if (h_stage) pinned_free(h_stage);
        if (okc) {
            float ms = 0;
            for (int q = 0; q < 3; ++q) { cudaEventElapsedTime(&ms, ev[q], ev[q + 1]); B->phase_ms[q] += ms; }
            cudaEventElapsedTime(&ms, ev[1], ev[4]); B->phase_ms[3] += ms;      // one-thread segments + unstuff
            cudaEventElapsedTime(&ms, ev[4], ev[5]); B->phase_ms[4] += ms;      // sync + repair + check
            cudaEventElapsedTime(&ms, ev[5], ev[2]); B->phase_ms[5] += ms;      // scan + write
        }
        for (auto& e : ev) cudaEventDestroy(e);
        if (!okc) { cudaStreamSynchronize(st); delete B; return nullptr; }     // nothing in flight may outlive the scratch it uses
        for (int k = 0; k < m; ++k) final_ok[live[li + k]] = status[k];
        li = lj;
    }
    for (int i = 0; i < n; ++i) {
        gb200_image_desc& D = B->images[i];
        if (!final_ok[i]) continue;
        const Parsed& p = P[i];
        int rc = req_comps_in < 0 ? p.comps : req_comps_in;
        D.status = 1; D.pixels = d_out + out_off[i];
        D.width = p.width; D.height = p.height; D.channels = rc; D.file_channels = p.comps; D.bits = 8;
        D.pixel_type = rc == 1 ? GB200_l8 : rc == 3 ? GB200_rgb8 : GB200_rgba8;       // plugins/jpeg.d:83-89
        D.pitch = p.width * rc;
        D.pixelAspectRatio = p.par; D.ppmY = p.ppiY; D.ppmX = p.ppiX;
    }
    B->device_ms = now_ms() - t0 - B->host_parse_ms;
    return B;
}

} // namespace gb

/* 0 = decodable here (baseline / extended sequential with one interleaved scan, or progressive SOF2), 2 = sequential but
 * non-interleaved multi-scan (not on this path), -1 = not a JPEG this parser accepts. (1 used to mean "progressive,
 * unsupported" and is no longer returned.) Header walk on the host, no GPU work. */
GB_API int gb200_jpeg_probe(const uint8_t* data, size_t len)
{
    Parsed pp;
    if (data && parse_jpeg(data, len, pp)) return 0;
    return pp.unsupported ? pp.unsupported : -1;
}

GB_API gb200_batch* gb200_jpeg_decode_batch(int n, const uint8_t* const* files, const size_t* lens,
                                            const uint8_t* const* files_dev, int req_comps, void* stream)
{
    gb::clear_error();
    return gb::jpeg_decode_batch(n, files, lens, files_dev, req_comps, (cudaStream_t)stream);
}

GB_API uint8_t* gb200_jpeg_load(const uint8_t* data, size_t len, int req_comps, int* width, int* height,
                                int* actual_comps, float* pixelAspectRatio, float* dotsPerInchY)
{
    gb::clear_error();
    if (pixelAspectRatio) *pixelAspectRatio = -1;
    if (dotsPerInchY) *dotsPerInchY = -1;
    if (!gb::ensure_device()) return nullptr;
    const uint8_t* f[1] = {data}; size_t l[1] = {len};
    cudaStream_t st = gb::thread_stream();
    gb200_batch* B = gb::jpeg_decode_batch(1, f, l, nullptr, req_comps, st);
    if (!B) return nullptr;
    const gb200_image_desc& D = B->images[0];
    if (!D.status) {
        // a valid file of a kind this path does not decode is reported as such, so that callers can route it elsewhere
        Parsed pp;
        parse_jpeg(data, len, pp);
        if (pp.unsupported == 2) gb::set_error("unsupported: non-interleaved multi-scan sequential JPEG -- the GPU path decodes single-scan interleaved files only");
        else gb::set_error("JPEG decoding failed");
        delete B; return nullptr;
    }
    size_t bytes = (size_t)D.pitch * D.height;
    uint8_t* out = (uint8_t*)malloc(bytes ? bytes : 1);
    if (!out) { delete B; return nullptr; }
    bool ok = gb::cuda_ok(cudaMemcpyAsync(out, D.pixels, bytes, cudaMemcpyDeviceToHost, st), "d2h", __FILE__, __LINE__) &&
              gb::cuda_ok(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__);
    if (width) *width = D.width; if (height) *height = D.height; if (actual_comps) *actual_comps = D.file_channels;
    if (pixelAspectRatio) *pixelAspectRatio = D.pixelAspectRatio;
    if (dotsPerInchY) *dotsPerInchY = D.ppmY;
    delete B;
    if (!ok) { free(out); return nullptr; }
    return out;
}
