// bmp.cu -- BMP decode (the first of the "other decoders") and format detection.
//
// stbi__bmp_load (codecs/stbdec.d:2263-2466, the stb_image 2.29 loader behind plugins/bmp.d:93-163) reads a header, an
// optional palette and then one pixel after the other from a byte stream. Everything that can fail -- and everything that
// depends on the stream's 128-byte refill buffer (a negative stbi__skip jumps to the end of that buffer, stbdec.d:822-
// 842) -- is the header walk, which stays on the host. The pixel loop has no dependency between pixels except the
// OR of all alpha values ("an alpha channel that is zero everywhere is replaced by 255", :2438-2443), so it is one
// thread per pixel: source bytes addressed from the pixel's coordinates (reads past the end of the file yield 0 like
// stbi__get8 at EOF), palette / bit-field extraction / BGR swap, the vertical flip folded into the destination
// address, a warp-reduced atomicOr for the alpha flag, and a second pass only where that flag or a conversion to 1 or
// 2 components (stbi__convert_format, :916-1054) needs one.
#include "common.h"
#include "batch.h"
#include <vector>
#include <chrono>

namespace {

// ---- stbi__context over a memory "callback" stream (stbdec.d:461-503, 780-842), host side ----
struct StbStream {
    const uint8_t* data; size_t len;
    size_t stream_pos = 0;
    uint8_t buf[128]; int buf_n = 0, buf_cur = 0;
    bool from_callbacks = true; int already_read = 0;
    StbStream(const uint8_t* d, size_t l) : data(d), len(l) { refill(); }
    void refill()
    {
        size_t n = len - stream_pos; if (n > 128) n = 128;
        if (n) memcpy(buf, data + stream_pos, n);
        stream_pos += n;
        already_read += buf_cur;
        if (n == 0) { from_callbacks = false; buf_cur = 0; buf_n = 1; buf[0] = 0; }
        else { buf_cur = 0; buf_n = (int)n; }
    }
    int get8() { if (buf_cur < buf_n) return buf[buf_cur++]; if (from_callbacks) { refill(); return buf[buf_cur++]; } return 0; }
    int get16le() { int z = get8(); return z + (get8() << 8); }
    uint32_t get32le() { uint32_t z = (uint32_t)get16le(); z += (uint32_t)get16le() << 16; return z; }
    void skip(int n)
    {
        if (n == 0) return;
        if (n < 0) { buf_cur = buf_n; return; }
        const int blen = buf_n - buf_cur;
        if (blen < n) { buf_cur = buf_n; const size_t adv = (size_t)(n - blen); stream_pos = stream_pos + adv > len ? len : stream_pos + adv; return; }
        buf_cur += n;
    }
};

struct BmpPlan {
    bool ok = false;
    int w = 0, h = 0, bpp = 0, img_n = 0, target = 0, req = 0, flip = 0, easy = 0;
    uint32_t mr = 0, mg = 0, mb = 0, ma = 0, all_a0 = 255;
    int rshift = 0, gshift = 0, bshift = 0, ashift = 0, rcount = 0, gcount = 0, bcount = 0, acount = 0;
    int row_bytes = 0;              // source bytes per row including the padding
    size_t data_off = 0;            // file offset of the first pixel byte
    uint8_t pal[256][4];
    float ppmX = -1, ppmY = -1, par = -1;
};

int high_bit(uint32_t z)            // stbi__high_bit :2468
{
    int n = 0;
    if (z == 0) return -1;
    if (z >= 0x10000) { n += 16; z >>= 16; }
    if (z >= 0x00100) { n += 8; z >>= 8; }
    if (z >= 0x00010) { n += 4; z >>= 4; }
    if (z >= 0x00004) { n += 2; z >>= 2; }
    if (z >= 0x00002) { n += 1; }
    return n;
}

// stbi__bmp_test + stbi__bmp_parse_header + the preamble of stbi__bmp_load (:2147-2340, :2381-2408)
bool plan_bmp(const uint8_t* data, size_t len, int req_comp, BmpPlan& P)
{
    {   // stbi__bmp_test_raw (:2241-2254) on the first buffer
        StbStream t(data, len);
        if (t.get8() != 'B' || t.get8() != 'M') return false;
        t.get32le(); t.get16le(); t.get16le(); t.get32le();
        const int sz = (int)t.get32le();
        if (!(sz == 12 || sz == 40 || sz == 56 || sz == 108 || sz == 124)) return false;
    }
    StbStream s(data, len);
    memset(P.pal, 0, sizeof(P.pal));
    int extra_read = 14;
    if (s.get8() != 'B' || s.get8() != 'M') return false;
    s.get32le(); s.get16le(); s.get16le();
    const int offset = (int)s.get32le();
    const int hsz = (int)s.get32le();
    uint32_t img_x, img_y;
    if (offset < 0) return false;
    if (hsz != 12 && hsz != 40 && hsz != 56 && hsz != 108 && hsz != 124) return false;
    if (hsz == 12) { img_x = (uint32_t)s.get16le(); img_y = (uint32_t)s.get16le(); }
    else { img_x = s.get32le(); img_y = s.get32le(); }
    if (s.get16le() != 1) return false;
    P.bpp = s.get16le();
    auto mask_defaults = [&](int compress) {             // stbi__bmp_set_mask_defaults :2121
        if (compress == 3) return;
        if (compress == 0) {
            if (P.bpp == 16) { P.mr = 31u << 10; P.mg = 31u << 5; P.mb = 31u; }
            else if (P.bpp == 32) { P.mr = 0xffu << 16; P.mg = 0xffu << 8; P.mb = 0xffu; P.ma = 0xffu << 24; P.all_a0 = 0; }
            else P.mr = P.mg = P.mb = P.ma = 0;
        }
    };
    if (hsz != 12) {
        const int compress = (int)s.get32le();
        if (compress == 1 || compress == 2) return false;
        if (compress >= 4) return false;
        if (compress == 3 && P.bpp != 16 && P.bpp != 32) return false;
        s.get32le();
        const int xppm = (int)s.get32le(), yppm = (int)s.get32le();
        if (xppm > 1) P.ppmX = (float)xppm;
        if (yppm > 1) P.ppmY = (float)yppm;
        if (P.ppmX != -1 && P.ppmY != -1) P.par = P.ppmX / P.ppmY;
        s.get32le(); s.get32le();
        if (hsz == 40 || hsz == 56) {
            if (hsz == 56) { s.get32le(); s.get32le(); s.get32le(); s.get32le(); }
            if (P.bpp == 16 || P.bpp == 32) {
                if (compress == 0) mask_defaults(compress);
                else if (compress == 3) {
                    P.mr = s.get32le(); P.mg = s.get32le(); P.mb = s.get32le();
                    extra_read += 12;
                    if (P.mr == P.mg && P.mg == P.mb) return false;
                } else return false;
            }
        } else {
            P.mr = s.get32le(); P.mg = s.get32le(); P.mb = s.get32le(); P.ma = s.get32le();
            if (compress != 3) mask_defaults(compress);
            s.get32le();
            for (int i = 0; i < 12; ++i) s.get32le();
            if (hsz == 124) { s.get32le(); s.get32le(); s.get32le(); s.get32le(); }
        }
    }
    // stbi__bmp_load proper
    P.flip = ((int)img_y) > 0;
    { const int iy = (int)img_y; img_y = (uint32_t)(iy < 0 ? -iy : iy); }
    if (img_y > (1u << 24) || img_x > (1u << 24)) return false;
    int psize = 0;
    if (hsz == 12) { if (P.bpp < 24) psize = (offset - extra_read - 24) / 3; }
    else { if (P.bpp < 16) psize = (offset - extra_read - hsz) >> 2; }
    if (psize == 0) {
        const int so_far = s.already_read + s.buf_cur;
        if (so_far <= 0 || so_far > 1024) return false;
        if (offset < so_far || offset - so_far > 256 * 4) return false;
        s.skip(offset - so_far);
    }
    if (P.bpp == 24 && P.ma == 0xff000000u) P.img_n = 3; else P.img_n = P.ma ? 4 : 3;
    P.req = req_comp;
    P.target = (req_comp && req_comp >= 3) ? req_comp : P.img_n;
    {   // stbi__mad3sizes_valid(target, x, y, 0) (:558)
        const long long a = P.target, b = (int)img_x, c = (int)img_y;
        if ((int)img_x < 0 || (int)img_y < 0) return false;
        if (b && a > 0x7fffffffLL / b) return false;
        if (c && a * b > 0x7fffffffLL / c) return false;
    }
    P.w = (int)img_x; P.h = (int)img_y;
    int width;
    if (P.bpp < 16) {
        if (psize == 0 || psize > 256) return false;
        for (int i = 0; i < psize; ++i) {
            P.pal[i][2] = (uint8_t)s.get8(); P.pal[i][1] = (uint8_t)s.get8(); P.pal[i][0] = (uint8_t)s.get8();
            if (hsz != 12) s.get8();
            P.pal[i][3] = 255;
        }
        s.skip(offset - extra_read - hsz - psize * (hsz == 12 ? 3 : 4));
        if (P.bpp == 1) width = (P.w + 7) >> 3;
        else if (P.bpp == 4) width = (P.w + 1) >> 1;
        else if (P.bpp == 8) width = P.w;
        else return false;
    } else {
        s.skip(offset - extra_read - hsz);
        if (P.bpp == 24) width = 3 * P.w;
        else if (P.bpp == 16) width = 2 * P.w;
        else width = 0;
        if (P.bpp == 24) P.easy = 1;
        else if (P.bpp == 32) { if (P.mb == 0xff && P.mg == 0xff00 && P.mr == 0x00ff0000 && P.ma == 0xff000000u) P.easy = 2; }
        if (!P.easy) {
            if (!P.mr || !P.mg || !P.mb) return false;
            P.rshift = high_bit(P.mr) - 7; P.rcount = __builtin_popcount(P.mr);
            P.gshift = high_bit(P.mg) - 7; P.gcount = __builtin_popcount(P.mg);
            P.bshift = high_bit(P.mb) - 7; P.bcount = __builtin_popcount(P.mb);
            P.ashift = high_bit(P.ma) - 7; P.acount = __builtin_popcount(P.ma);
            if (P.rcount > 8 || P.gcount > 8 || P.bcount > 8 || P.acount > 8) return false;
        }
    }
    const int pad = (-width) & 3;
    if (P.bpp < 16) P.row_bytes = width + pad;
    else if (P.bpp == 24) P.row_bytes = 3 * P.w + pad;
    else if (P.bpp == 16) P.row_bytes = 2 * P.w + pad;
    else P.row_bytes = 4 * P.w;                       // 32-bit pixels, pad = 0 (other depths cannot get here: their masks are zero)
    // file offset of the next byte the stream would deliver; at or past the end everything reads as 0 (stbi__get8 :797)
    P.data_off = s.from_callbacks ? s.stream_pos - (size_t)(s.buf_n - s.buf_cur) : len;
    P.ok = true;
    return true;
}

struct BmpJob {
    const uint8_t* data; uint32_t len, data_off;
    int w, h, bpp, target, flip, easy, row_bytes;
    uint32_t mr, mg, mb, ma;
    int rshift, gshift, bshift, ashift, rcount, gcount, bcount, acount;
    uint8_t* out;                   // target components per pixel
    uint32_t* all_a;                // OR of every alpha value (starts at the header's value)
    uint32_t pix_base;              // first global pixel index of this image (prefix over the batch)
    uint8_t pal[256][4];
};
struct BmpFix {                     // second pass: alpha replacement and / or conversion to 1 or 2 components
    const uint8_t* src; uint8_t* dst; int w, h, target, req; const uint32_t* all_a; uint32_t pix_base;
};

__device__ __forceinline__ uint32_t rd8(const BmpJob& J, uint32_t off) { return off < J.len ? J.data[off] : 0u; }

__device__ __forceinline__ int bmp_shiftsigned(uint32_t v, int shift, int bits)       // stbi__shiftsigned :2493
{
    const uint32_t mul_table[9] = {0, 0xff, 0x55, 0x49, 0x11, 0x21, 0x41, 0x81, 0x01};
    const uint32_t shift_table[9] = {0, 0, 0, 1, 0, 2, 4, 6, 0};
    if (shift < 0) v <<= -shift; else v >>= shift;
    v >>= (8 - bits);
    return (int)((uint32_t)v * mul_table[bits]) >> shift_table[bits];
}

template <class J>
__device__ __forceinline__ int find_job(const J* jobs, int n, uint32_t g)
{
    int lo = 0, hi = n - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (jobs[mid].pix_base <= g) lo = mid; else hi = mid - 1; }
    return lo;
}

__global__ void __launch_bounds__(256)
bmp_decode_kernel(const BmpJob* __restrict__ jobs, int njobs, uint32_t total_pixels)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = g < total_pixels;
    const int ji = find_job(jobs, njobs, live ? g : total_pixels - 1);
    const BmpJob& J = jobs[ji];
    uint32_t a = 0;
    bool want_a = false;
    if (live) {
        const uint32_t lp = g - J.pix_base;
        const uint32_t j = lp / (uint32_t)J.w, i = lp - j * (uint32_t)J.w;
        const uint32_t base = J.data_off + j * (uint32_t)J.row_bytes;
        uint32_t r, gg, b;
        a = 255;
        if (J.bpp < 16) {
            uint32_t v;
            if (J.bpp == 1) v = (rd8(J, base + (i >> 3)) >> (7 - (i & 7))) & 1u;
            else if (J.bpp == 4) { const uint32_t x = rd8(J, base + (i >> 1)); v = (i & 1) ? (x & 15u) : (x >> 4); }
            else v = rd8(J, base + i);
            r = J.pal[v][0]; gg = J.pal[v][1]; b = J.pal[v][2];
        } else if (J.easy) {
            const uint32_t o = base + i * (J.easy == 2 ? 4u : 3u);
            b = rd8(J, o); gg = rd8(J, o + 1); r = rd8(J, o + 2);
            if (J.easy == 2) a = rd8(J, o + 3);
            want_a = true;
        } else {
            uint32_t v;
            if (J.bpp == 16) { const uint32_t o = base + i * 2u; v = rd8(J, o) | (rd8(J, o + 1) << 8); }
            else { const uint32_t o = base + i * 4u; v = rd8(J, o) | (rd8(J, o + 1) << 8) | (rd8(J, o + 2) << 16) | (rd8(J, o + 3) << 24); }
            r = (uint32_t)bmp_shiftsigned(v & J.mr, J.rshift, J.rcount) & 255u;
            gg = (uint32_t)bmp_shiftsigned(v & J.mg, J.gshift, J.gcount) & 255u;
            b = (uint32_t)bmp_shiftsigned(v & J.mb, J.bshift, J.bcount) & 255u;
            a = J.ma ? (uint32_t)bmp_shiftsigned(v & J.ma, J.ashift, J.acount) : 255u;
            want_a = true;
        }
        const uint32_t row = J.flip ? (uint32_t)J.h - 1u - j : j;
        uint8_t* o = J.out + ((size_t)row * J.w + i) * J.target;
        o[0] = (uint8_t)r; o[1] = (uint8_t)gg; o[2] = (uint8_t)b;
        if (J.target == 4) o[3] = (uint8_t)a;
    }
    // all_a |= a (:2418, :2430): one atomic per warp and image. Lanes of a warp may belong to two images at a boundary.
    if (__any_sync(0xffffffffu, want_a)) {
        const int j0 = __shfl_sync(0xffffffffu, ji, 0);
        const bool same = __all_sync(0xffffffffu, ji == j0);
        if (same) {
            const uint32_t m = __reduce_or_sync(0xffffffffu, (live && want_a) ? a : 0u);
            if ((threadIdx.x & 31) == 0 && m) atomicOr(jobs[j0].all_a, m);
        } else if (live && want_a && a) atomicOr(J.all_a, a);
    }
}

// The two "easy" layouts of stbi__bmp_load (:2396-2420: 24-bit BGR and 32-bit BGRA with the default masks) are a
// byte shuffle with a vertical flip, i.e. a copy: four pixels per thread, the source read as aligned 32-bit words
// realigned with a funnel shift (rows start at data_off + j * row_bytes, usually 2 mod 4), B and R swapped with one
// PRMT per pixel, whole words stored. A warp reads and writes 384 (or 512) contiguous bytes per instruction.
// grid = (ceil(w / 1024), h, images); SB = source bytes per pixel, DB = destination bytes per pixel.
template <int SB, int DB>
__global__ void __launch_bounds__(256)
bmp_easy_kernel(const BmpJob* __restrict__ jobs)
{
    const BmpJob& J = jobs[blockIdx.z];
    const uint32_t j = blockIdx.y;
    if (j >= (uint32_t)J.h) return;
    const uint32_t i0 = (blockIdx.x * 256u + threadIdx.x) * 4u;
    const bool live = i0 < (uint32_t)J.w;
    uint32_t aor = 0;
    if (live) {
        const uint32_t npx = min(4u, (uint32_t)J.w - i0);
        const uint32_t o = J.data_off + j * (uint32_t)J.row_bytes + i0 * SB;
        uint32_t px[4];                                  // B | G << 8 | R << 16 | A << 24
        if ((uint64_t)o + 4 * SB + 4 <= J.len) {
            const uint8_t* p = J.data + o;
            const uint32_t mis = (uint32_t)((uintptr_t)p & 3u), sh = mis * 8u;
            const uint32_t* w = (const uint32_t*)(p - mis);
            uint32_t x[SB + 1], v[SB];
#pragma unroll
            for (int k = 0; k <= SB; ++k) x[k] = __ldg(w + k);
#pragma unroll
            for (int k = 0; k < SB; ++k) v[k] = __funnelshift_r(x[k], x[k + 1], sh);
            if (SB == 3) {
                px[0] = v[0] & 0x00ffffffu; px[1] = __funnelshift_r(v[0], v[1], 24) & 0x00ffffffu;
                px[2] = __funnelshift_r(v[1], v[2], 16) & 0x00ffffffu; px[3] = v[2] >> 8;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) px[k] = v[k % SB];
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                px[k] = 0;
#pragma unroll
                for (int c = 0; c < SB; ++c) px[k] |= rd8(J, o + k * SB + c) << (8 * c);
            }
        }
        uint32_t q[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t a = SB == 4 ? px[k] >> 24 : 255u;
            if (k < (int)npx) aor |= a;
            q[k] = (__byte_perm(px[k], 0, 0x3012) & 0x00ffffffu) | (a << 24);       // R | G << 8 | B << 16 | A << 24
        }
        const uint32_t row = J.flip ? (uint32_t)J.h - 1u - j : j;
        uint8_t* d = J.out + ((size_t)row * J.w + i0) * DB;
        if (DB == 4) {
            if (npx == 4 && (((uintptr_t)d) & 15) == 0) *(uint4*)d = make_uint4(q[0], q[1], q[2], q[3]);
            else for (uint32_t k = 0; k < npx; ++k) ((uint32_t*)d)[k] = q[k];
        } else {
            if (npx == 4 && (((uintptr_t)d) & 3) == 0) {
                uint32_t* d32 = (uint32_t*)d;
                d32[0] = (q[0] & 0x00ffffffu) | (q[1] << 24);
                d32[1] = ((q[1] >> 8) & 0x0000ffffu) | (q[2] << 16);
                d32[2] = ((q[2] >> 16) & 0x000000ffu) | (q[3] << 8);
            } else {
                for (uint32_t k = 0; k < npx; ++k) { d[k * 3] = (uint8_t)q[k]; d[k * 3 + 1] = (uint8_t)(q[k] >> 8); d[k * 3 + 2] = (uint8_t)(q[k] >> 16); }
            }
        }
    }
    if (SB == 4) {                                       // all_a |= a (:2418)
        const uint32_t m = __reduce_or_sync(0xffffffffu, aor);
        if ((threadIdx.x & 31) == 0 && m) atomicOr(J.all_a, m);
    }
}

__device__ __forceinline__ uint8_t bmp_compute_y(int r, int g, int b) { return (uint8_t)(((r * 77) + (g * 150) + (29 * b)) >> 8); }   // :911

__global__ void __launch_bounds__(256)
bmp_fix_kernel(const BmpFix* __restrict__ jobs, int njobs, uint32_t total_pixels)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_pixels) return;
    const BmpFix& J = jobs[find_job(jobs, njobs, g)];
    const uint32_t lp = g - J.pix_base;
    const uint8_t* s = J.src + (size_t)lp * J.target;
    uint32_t a = J.target == 4 ? s[3] : 255u;
    if (J.target == 4 && *J.all_a == 0) a = 255;                  // "if alpha channel is all 0s, replace with all 255s" (:2438)
    if (J.req == J.target) { if (J.target == 4) J.dst[(size_t)lp * 4 + 3] = (uint8_t)a; return; }
    uint8_t* d = J.dst + (size_t)lp * J.req;                      // stbi__convert_format, 3|4 -> 1|2 (:916-1054)
    d[0] = bmp_compute_y(s[0], s[1], s[2]);
    if (J.req == 2) d[1] = (uint8_t)a;
}

inline size_t al(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
inline double now_ms() { using namespace std::chrono; return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count(); }

} // namespace

namespace gb {

gb200_batch* bmp_decode_batch(int n, const uint8_t* const* files, const size_t* lens, const uint8_t* const* files_dev,
                              int req_comp, cudaStream_t st)
{
    if (!ensure_device()) return nullptr;
    if (n < 0) { set_error("bmp_decode_batch: negative count"); return nullptr; }
    gb200_batch* B = new gb200_batch;
    B->stream = st;
    B->images.resize((size_t)n);
    for (auto& D : B->images) { memset(&D, 0, sizeof(D)); D.ppmX = D.ppmY = D.pixelAspectRatio = -1; }
    if (req_comp < 0 || req_comp > 4) return B;
    const double t0 = now_ms();
    std::vector<BmpPlan> P((size_t)n);
    std::vector<int> live;
    size_t out_total = 0, tmp_total = 0, file_total = 0;
    std::vector<size_t> out_off((size_t)n, 0), tmp_off((size_t)n, 0), file_off((size_t)n, 0);
    uint64_t total_pixels = 0;
    for (int i = 0; i < n; ++i) {
        if (!files[i] || lens[i] > 0xfffffff0u || !plan_bmp(files[i], lens[i], req_comp, P[i])) continue;
        const uint64_t px = (uint64_t)P[i].w * P[i].h;
        if (px == 0 || total_pixels + px > 0xfffffff0ull) { P[i].ok = false; continue; }
        live.push_back(i);
        total_pixels += px;
        const int outc = req_comp ? req_comp : P[i].img_n;
        out_off[i] = out_total; out_total += al((size_t)px * outc);
        if (outc != P[i].target) { tmp_off[i] = tmp_total; tmp_total += al((size_t)px * P[i].target); }
        file_off[i] = file_total; file_total += al(lens[i] + 16);
    }
    B->host_parse_ms = now_ms() - t0;
    const int m = (int)live.size();
    if (!m) return B;
    uint8_t* d_out = (uint8_t*)dev_alloc(out_total);
    if (!d_out) { delete B; return nullptr; }
    B->device_allocs.push_back(d_out);
    DevBuf d_files(files_dev ? 256 : file_total), d_tmp(tmp_total + 256), d_jobs(sizeof(BmpJob) * (size_t)m), d_fix(sizeof(BmpFix) * (size_t)m),
           d_flags(4 * (size_t)m);
    if (!d_files.p || !d_tmp.p || !d_jobs.p || !d_fix.p || !d_flags.p) { delete B; return nullptr; }
    uint8_t* h_stage = nullptr;
    if (!files_dev) { h_stage = (uint8_t*)pinned_alloc(file_total); if (!h_stage) { delete B; return nullptr; } }
    // jobs[0 .. ngen) go to the per-pixel kernel, then four groups for the easy layouts (source 3|4 bytes x target 3|4)
    std::vector<BmpJob> jobs((size_t)m); std::vector<BmpFix> fix; std::vector<uint32_t> flags((size_t)m);
    std::vector<HostCopy> hcopies;
    auto group_of = [&](const BmpPlan& p) { return p.easy && p.h <= 65535 && (p.target == 3 || p.target == 4) ? 1 + (p.easy - 1) * 2 + (p.target - 3) : 0; };
    int gcount[5] = {0, 0, 0, 0, 0}, gfirst[6] = {0, 0, 0, 0, 0, 0}, gmaxw[5] = {0, 0, 0, 0, 0}, gmaxh[5] = {0, 0, 0, 0, 0};
    for (int i : live) { const int g = group_of(P[i]); ++gcount[g]; gmaxw[g] = std::max(gmaxw[g], P[i].w); gmaxh[g] = std::max(gmaxh[g], P[i].h); }
    for (int g = 0; g < 5; ++g) gfirst[g + 1] = gfirst[g] + gcount[g];
    int gnext[5]; for (int g = 0; g < 5; ++g) gnext[g] = gfirst[g];
    uint32_t pix = 0, fixpix = 0;
    for (int i : live) {
        const BmpPlan& p = P[i];
        const int g = group_of(p);
        const int k = gnext[g]++;
        BmpJob& J = jobs[k];
        memset(&J, 0, sizeof(J));
        if (files_dev) J.data = files_dev[i];
        else { hcopies.push_back(HostCopy{h_stage + file_off[i], files[i], lens[i]}); J.data = d_files.as<uint8_t>() + file_off[i]; }
        J.len = (uint32_t)lens[i]; J.data_off = (uint32_t)std::min<size_t>(p.data_off, lens[i]);
        J.w = p.w; J.h = p.h; J.bpp = p.bpp; J.target = p.target; J.flip = p.flip; J.easy = p.easy; J.row_bytes = p.row_bytes;
        J.mr = p.mr; J.mg = p.mg; J.mb = p.mb; J.ma = p.ma;
        J.rshift = p.rshift; J.gshift = p.gshift; J.bshift = p.bshift; J.ashift = p.ashift;
        J.rcount = p.rcount; J.gcount = p.gcount; J.bcount = p.bcount; J.acount = p.acount;
        memcpy(J.pal, p.pal, sizeof(J.pal));
        const int outc = req_comp ? req_comp : p.img_n;
        const bool convert = outc != p.target;
        J.out = convert ? d_tmp.as<uint8_t>() + tmp_off[i] : d_out + out_off[i];
        J.all_a = d_flags.as<uint32_t>() + k;
        flags[k] = p.all_a0;
        if (g == 0) { J.pix_base = pix; pix += (uint32_t)p.w * (uint32_t)p.h; }
        if (convert || (p.target == 4 && p.all_a0 == 0)) {
            BmpFix F; F.src = J.out; F.dst = d_out + out_off[i]; F.w = p.w; F.h = p.h; F.target = p.target; F.req = outc;
            F.all_a = J.all_a; F.pix_base = fixpix;
            fixpix += (uint32_t)p.w * (uint32_t)p.h;
            fix.push_back(F);
        }
    }
    host_copy_parallel(hcopies.data(), hcopies.size());
    cudaEvent_t ev[3];
    for (auto& e : ev) cudaEventCreate(&e);
    bool okc = true;
    cudaEventRecord(ev[0], st);
    if (!files_dev) okc &= cuda_ok(cudaMemcpyAsync(d_files.p, h_stage, file_total, cudaMemcpyHostToDevice, st), "files", __FILE__, __LINE__);
    okc &= cuda_ok(cudaMemcpyAsync(d_jobs.p, jobs.data(), sizeof(BmpJob) * m, cudaMemcpyHostToDevice, st), "jobs", __FILE__, __LINE__);
    okc &= cuda_ok(cudaMemcpyAsync(d_flags.p, flags.data(), 4 * (size_t)m, cudaMemcpyHostToDevice, st), "flags", __FILE__, __LINE__);
    if (!fix.empty()) okc &= cuda_ok(cudaMemcpyAsync(d_fix.p, fix.data(), sizeof(BmpFix) * fix.size(), cudaMemcpyHostToDevice, st), "fix", __FILE__, __LINE__);
    cudaEventRecord(ev[1], st);
    if (okc) {
        if (gcount[0]) { bmp_decode_kernel<<<(pix + 255) / 256, 256, 0, st>>>(d_jobs.as<BmpJob>(), gcount[0], pix); count_launch(); }
        for (int g = 1; g < 5; ++g) {
            for (int z0 = 0; z0 < gcount[g]; z0 += 65535) {
                const dim3 grid((unsigned)((gmaxw[g] + 1023) / 1024), (unsigned)gmaxh[g], (unsigned)std::min(65535, gcount[g] - z0));
                const BmpJob* dj = d_jobs.as<BmpJob>() + gfirst[g] + z0;
                if (g == 1) bmp_easy_kernel<3, 3><<<grid, 256, 0, st>>>(dj);
                else if (g == 2) bmp_easy_kernel<3, 4><<<grid, 256, 0, st>>>(dj);
                else if (g == 3) bmp_easy_kernel<4, 3><<<grid, 256, 0, st>>>(dj);
                else bmp_easy_kernel<4, 4><<<grid, 256, 0, st>>>(dj);
                count_launch();
            }
        }
        if (!fix.empty()) { bmp_fix_kernel<<<(fixpix + 255) / 256, 256, 0, st>>>(d_fix.as<BmpFix>(), (int)fix.size(), fixpix); count_launch(); }
    }
    cudaEventRecord(ev[2], st);
    okc &= cuda_ok(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__);
    okc &= cuda_ok(cudaGetLastError(), "kernels", __FILE__, __LINE__);
    if (h_stage) pinned_free(h_stage);
    if (okc) for (int q = 0; q < 2; ++q) { float ms = 0; cudaEventElapsedTime(&ms, ev[q], ev[q + 1]); B->phase_ms[q] += ms; }
    for (auto& e : ev) cudaEventDestroy(e);
    if (!okc) { cudaStreamSynchronize(st); delete B; return nullptr; }
    for (int i : live) {
        gb200_image_desc& D = B->images[i];
        const BmpPlan& p = P[i];
        const int outc = req_comp ? req_comp : p.img_n;
        D.status = 1; D.pixels = d_out + out_off[i];
        D.width = p.w; D.height = p.h; D.channels = outc; D.file_channels = p.img_n; D.bits = 8;
        D.pixel_type = outc == 1 ? GB200_l8 : outc == 2 ? GB200_la8 : outc == 3 ? GB200_rgb8 : GB200_rgba8;      // plugins/bmp.d:138-155
        D.pitch = p.w * outc;
        D.ppmX = p.ppmX; D.ppmY = p.ppmY; D.pixelAspectRatio = p.par;
    }
    B->device_ms = now_ms() - t0 - B->host_parse_ms;
    return B;
}

} // namespace gb

GB_API gb200_batch* gb200_bmp_decode_batch(int n, const uint8_t* const* files, const size_t* lens,
                                           const uint8_t* const* files_dev, int req_comp, void* stream)
{
    gb::clear_error();
    return gb::bmp_decode_batch(n, files, lens, files_dev, req_comp, (cudaStream_t)stream);
}

// stbi_load_from_callbacks on a BMP (stbdec.d:725 -> :2263), as loadBMP calls it (plugins/bmp.d:112): host bytes in,
// malloc'd host pixels out, *comp = channels in the file.
GB_API uint8_t* gb200_bmp_load(const uint8_t* data, size_t len, int req_comp, int* width, int* height, int* comp,
                               float* ppmX, float* ppmY, float* pixelRatio)
{
    gb::clear_error();
    if (!gb::ensure_device()) return nullptr;
    const uint8_t* f[1] = {data}; size_t l[1] = {len};
    cudaStream_t st = gb::thread_stream();
    gb200_batch* B = gb::bmp_decode_batch(1, f, l, nullptr, req_comp, st);
    if (!B) return nullptr;
    const gb200_image_desc& D = B->images[0];
    if (!D.status) { gb::set_error("BMP decoding failed"); delete B; return nullptr; }
    const size_t bytes = (size_t)D.pitch * D.height;
    uint8_t* out = (uint8_t*)malloc(bytes ? bytes : 1);
    const bool ok = out && gb::cuda_ok(cudaMemcpyAsync(out, D.pixels, bytes, cudaMemcpyDeviceToHost, st), "d2h", __FILE__, __LINE__) &&
                    gb::cuda_ok(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__);
    if (width) *width = D.width;
    if (height) *height = D.height;
    if (comp) *comp = D.file_channels;
    if (ppmX) *ppmX = D.ppmX;
    if (ppmY) *ppmY = D.ppmY;
    if (pixelRatio) *pixelRatio = D.pixelAspectRatio;
    delete B;
    if (!ok) { free(out); return nullptr; }
    return out;
}

// Image.identifyFormatFromMemory (image.d:1037-1061): the detect procs of every plugin in ImageFormat order, TGA last
// (its test is fuzzy). Host only. Returns the ImageFormat value (types.d:14-28) or -1 (unknown).
GB_API int gb200_identify_format(const uint8_t* d, size_t len)
{
    auto starts = [&](const char* sig, size_t n) { return d && len >= n && memcmp(d, sig, n) == 0; };
    if (starts("\xff\xd8", 2)) return GB200_FORMAT_JPEG;                          // plugins/jpeg.d:106
    if (starts("\x89PNG\r\n\x1a\n", 8)) return GB200_FORMAT_PNG;                  // png.d:165
    if (starts("qoif", 4)) return GB200_FORMAT_QOI;                               // qoi.d:143
    if (starts("qoix", 4)) return GB200_FORMAT_QOIX;                              // qoix.d:149
    if (starts("DDS ", 4)) return GB200_FORMAT_DDS;                               // dds.d:40
    if (starts("GIF87a", 6) || starts("GIF89a", 6)) return GB200_FORMAT_GIF;      // gif.d:42
    if (d && len >= 18 && d[0] == 'B' && d[1] == 'M') {                           // bmp.d:45 (accepts 52, which the loader rejects)
        const uint32_t ds = d[14] | (d[15] << 8) | (d[16] << 16) | ((uint32_t)d[17] << 24);
        if (ds == 12 || ds == 40 || ds == 52 || ds == 56 || ds == 108 || ds == 124) return GB200_FORMAT_BMP;
    }
    if (starts("\xff\x0a", 2)) return GB200_FORMAT_JXL;                           // jxl.d:142
    if (starts("\xa5", 1)) return GB200_FORMAT_SQZ;                               // sqz.d:135
    {   // tga.d:97 -> TGADecoder.getImageInfo (codecs/tga.d:313-382)
        size_t p = 0;
        auto r8 = [&](int& v) { if (!d || p >= len) return false; v = d[p++]; return true; };
        auto r16 = [&](int& v) { if (!d || p + 2 > len) return false; v = d[p] | (d[p + 1] << 8); p += 2; return true; };
        auto sk = [&](size_t n) { if (p + n > len) return false; p += n; return true; };
        int off, cmap, type, ps, pl, cs, w, h, bpp;
        bool ok = r8(off) && r8(cmap) && cmap <= 1 && r8(type);
        if (ok) {
            if (cmap == 1) ok = (type == 1 || type == 9) && r16(ps) && r16(pl) && pl != 0 && r8(cs) && (cs == 8 || cs == 15 || cs == 16 || cs == 24 || cs == 32) && sk(4);
            else ok = (type == 2 || type == 3 || type == 10 || type == 11) && sk(9);
        }
        ok = ok && r16(w) && r16(h) && w >= 1 && h >= 1 && r8(bpp);
        ok = ok && !(cmap == 1 && bpp != 8 && bpp != 16) && (bpp == 8 || bpp == 15 || bpp == 16 || bpp == 24 || bpp == 32);
        if (ok) return GB200_FORMAT_TGA;
    }
    return -1;
}
