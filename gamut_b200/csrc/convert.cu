// convert.cu -- PixelType scanline converters for sm_100a.
//
// Replaces the row loops of source/gamut/scanline.d:70-930 (scanlinesConvert and the 46
// scanline_convert_* functions). The two reference stages (src -> intermediate -> dst, intermediate
// = rgba8 iff both types are in {l8,la8,rgb8,rgba8}, else rgbaf32; scanline.d:25-31) are fused in
// registers; the intermediate float32 values are materialised with the reference's rounding:
// every operation is a separate IEEE binary32 op in source order (__fdiv_rn/__fmul_rn/__fadd_rn, the
// TU is also compiled with -fmad=false), casts truncate toward zero with cvttss2si semantics.
//
// Two kernels:
//   convert_direct<S,D>  one lane per pixel, 4/8/16-byte vector load + vector store, fully coalesced;
//                        used for the pixel sizes that are powers of two >= 4 on both sides and aligned
//                        (config 2: rgba8 <-> rgbaf32).
//   convert_staged<D>    generic: a CTA stages a tile of source bytes in shared memory with 16-byte
//                        vector loads, converts pixel by pixel out of shared memory into a shared
//                        output tile, and writes that tile with 16-byte vector stores. Handles 1/2/3/6/12
//                        byte pixels, arbitrary (also negative) pitches and unaligned row starts while
//                        keeping HBM traffic coalesced.
// HBM-bound: algorithmic bytes/pixel = size(src)+size(dst) (20 B/px for rgba8<->rgbaf32).
#include "common.h"
#include "../../include/gamut_b200.h"

namespace {

struct F4 { float r, g, b, a; };

// v / 255.0f and v / 65535.0f (IEEE division in the reference, scanline.d:246,260,434...) without the division
// sequence: q0 = v * fl(1/D), r = fma(-q0, D, v) (exact residual), q = fma(r, fl(1/D), q0). For these divisors the
// result equals the correctly rounded quotient for EVERY integer numerator 0..255 / 0..65535 -- checked exhaustively
// in exact rational arithmetic and by tests/test_convert_gpu.py against the reference-text
// vectors. Three FP32 instructions instead of ~10; the forward rgba8 -> rgbaf32 kernel was issue-bound on them.
__device__ __forceinline__ float div_const(float v, float d, float rc)
{
    const float q0 = __fmul_rn(v, rc);
    const float r = __fmaf_rn(-q0, d, v);
    return __fmaf_rn(r, rc, q0);
}
__device__ __forceinline__ float n8(unsigned v)  { return div_const((float)v, 255.0f, __uint_as_float(0x3b808081u)); }
__device__ __forceinline__ float n16(unsigned v) { return div_const((float)v, 65535.0f, __uint_as_float(0x37800080u)); }

// cast(T)(float) as x86-64 cvttss2si + truncation: out-of-int32-range and NaN give 0x80000000.
__device__ __forceinline__ int cvtt(float t)
{
    int i = __float2int_rz(t);
    if (t >= 2147483648.0f) i = (int)0x80000000;
    return i;
}
__device__ __forceinline__ uint8_t  q8(float t)  { return (uint8_t)cvtt(__fadd_rn(0.5f, t)); }
__device__ __forceinline__ uint16_t q16(float t) { return (uint16_t)cvtt(__fadd_rn(0.5f, t)); }

__device__ __forceinline__ uint16_t ld16(const uint8_t* p) { return *(const uint16_t*)p; }
__device__ __forceinline__ float    ldf(const uint8_t* p)  { return *(const float*)p; }
__device__ __forceinline__ void st16(uint8_t* p, uint16_t v) { *(uint16_t*)p = v; }
__device__ __forceinline__ void stf(uint8_t* p, float v)     { *(float*)p = v; }

// ---- source pixel -> rgbaf32 (scanline.d:240-529). p is element-aligned. ----
__device__ __forceinline__ F4 unpremul(F4 v)
{
    if (v.a != 0.0f) { v.r = __fdiv_rn(v.r, v.a); v.g = __fdiv_rn(v.g, v.a); v.b = __fdiv_rn(v.b, v.a); }
    return v;
}
__device__ __forceinline__ F4 load_f32(int S, const uint8_t* p)
{
    F4 v;
    switch (S) {
    case GB200_l8:    { float b = n8(p[0]); v = {b, b, b, 1.0f}; } break;
    case GB200_l16:   { float b = n16(ld16(p)); v = {b, b, b, 1.0f}; } break;
    case GB200_lf32:  { float b = ldf(p); v = {b, b, b, 1.0f}; } break;
    case GB200_la8:   { float b = n8(p[0]); v = {b, b, b, n8(p[1])}; } break;
    case GB200_la16:  { float b = n16(ld16(p)); v = {b, b, b, n16(ld16(p + 2))}; } break;
    case GB200_laf32: { float b = ldf(p); v = {b, b, b, ldf(p + 4)}; } break;
    case GB200_lap8:  { float b = n8(p[0]), a = n8(p[1]); if (a != 0.0f) b = __fdiv_rn(b, a); v = {b, b, b, a}; } break;
    case GB200_lap16: { float b = n16(ld16(p)), a = n16(ld16(p + 2)); if (a != 0.0f) b = __fdiv_rn(b, a); v = {b, b, b, a}; } break;
    case GB200_lapf32:{ float b = ldf(p), a = ldf(p + 4); if (a != 0.0f) b = __fdiv_rn(b, a); v = {b, b, b, a}; } break;
    case GB200_rgb8:   v = {n8(p[0]), n8(p[1]), n8(p[2]), 1.0f}; break;
    case GB200_rgb16:  v = {n16(ld16(p)), n16(ld16(p + 2)), n16(ld16(p + 4)), 1.0f}; break;
    case GB200_rgbf32: v = {ldf(p), ldf(p + 4), ldf(p + 8), 1.0f}; break;
    case GB200_rgba8:  v = {n8(p[0]), n8(p[1]), n8(p[2]), n8(p[3])}; break;
    case GB200_rgba16: v = {n16(ld16(p)), n16(ld16(p + 2)), n16(ld16(p + 4)), n16(ld16(p + 6))}; break;
    case GB200_rgbaf32:v = {ldf(p), ldf(p + 4), ldf(p + 8), ldf(p + 12)}; break;
    case GB200_rgbap8: v = unpremul(F4{n8(p[0]), n8(p[1]), n8(p[2]), n8(p[3])}); break;
    case GB200_rgbap16:v = unpremul(F4{n16(ld16(p)), n16(ld16(p + 2)), n16(ld16(p + 4)), n16(ld16(p + 6))}); break;
    default:           v = unpremul(F4{ldf(p), ldf(p + 4), ldf(p + 8), ldf(p + 12)}); break; // rgbapf32
    }
    return v;
}

// ---- rgbaf32 -> destination pixel (scanline.d:539-803) ----
template <int D>
__device__ __forceinline__ void store_f32(F4 v, uint8_t* p)
{
    // grey = (r + g + b), left-assoc
    if (D == GB200_l8)    { float s = __fadd_rn(__fadd_rn(v.r, v.g), v.b); p[0] = q8(__fdiv_rn(__fmul_rn(s, 255.0f), 3.0f)); }
    if (D == GB200_l16)   { float s = __fadd_rn(__fadd_rn(v.r, v.g), v.b); st16(p, q16(__fdiv_rn(__fmul_rn(s, 65535.0f), 3.0f))); }
    if (D == GB200_lf32)  { float s = __fadd_rn(__fadd_rn(v.r, v.g), v.b); stf(p, __fdiv_rn(s, 3.0f)); }
    if (D == GB200_la8)   { float s = __fadd_rn(__fadd_rn(v.r, v.g), v.b); p[0] = q8(__fdiv_rn(__fmul_rn(s, 255.0f), 3.0f)); p[1] = q8(__fmul_rn(v.a, 255.0f)); }
    if (D == GB200_la16)  { float s = __fadd_rn(__fadd_rn(v.r, v.g), v.b); st16(p, q16(__fdiv_rn(__fmul_rn(s, 65535.0f), 3.0f))); st16(p + 2, q16(__fmul_rn(v.a, 65535.0f))); }
    if (D == GB200_laf32) { float s = __fadd_rn(__fadd_rn(v.r, v.g), v.b); stf(p, __fdiv_rn(s, 3.0f)); stf(p + 4, v.a); }
    if (D == GB200_lap8)  { float s = __fadd_rn(__fadd_rn(v.r, v.g), v.b); p[0] = q8(__fdiv_rn(__fmul_rn(__fmul_rn(s, v.a), 255.0f), 3.0f)); p[1] = q8(__fmul_rn(v.a, 255.0f)); }
    if (D == GB200_lap16) { float s = __fadd_rn(__fadd_rn(v.r, v.g), v.b); st16(p, q16(__fdiv_rn(__fmul_rn(__fmul_rn(s, v.a), 65535.0f), 3.0f))); st16(p + 2, q16(__fmul_rn(v.a, 65535.0f))); }
    if (D == GB200_lapf32){ float s = __fadd_rn(__fadd_rn(v.r, v.g), v.b); stf(p, __fdiv_rn(__fmul_rn(s, v.a), 3.0f)); stf(p + 4, v.a); }
    if (D == GB200_rgb8)  { p[0] = q8(__fmul_rn(v.r, 255.0f)); p[1] = q8(__fmul_rn(v.g, 255.0f)); p[2] = q8(__fmul_rn(v.b, 255.0f)); }
    if (D == GB200_rgb16) { st16(p, q16(__fmul_rn(v.r, 65535.0f))); st16(p + 2, q16(__fmul_rn(v.g, 65535.0f))); st16(p + 4, q16(__fmul_rn(v.b, 65535.0f))); }
    if (D == GB200_rgbf32){ stf(p, v.r); stf(p + 4, v.g); stf(p + 8, v.b); }
    if (D == GB200_rgba8) { p[0] = q8(__fmul_rn(v.r, 255.0f)); p[1] = q8(__fmul_rn(v.g, 255.0f)); p[2] = q8(__fmul_rn(v.b, 255.0f)); p[3] = q8(__fmul_rn(v.a, 255.0f)); }
    if (D == GB200_rgba16){ st16(p, q16(__fmul_rn(v.r, 65535.0f))); st16(p + 2, q16(__fmul_rn(v.g, 65535.0f))); st16(p + 4, q16(__fmul_rn(v.b, 65535.0f))); st16(p + 6, q16(__fmul_rn(v.a, 65535.0f))); }
    if (D == GB200_rgbaf32){ stf(p, v.r); stf(p + 4, v.g); stf(p + 8, v.b); stf(p + 12, v.a); }
    if (D == GB200_rgbap8) { p[0] = q8(__fmul_rn(__fmul_rn(v.r, v.a), 255.0f)); p[1] = q8(__fmul_rn(__fmul_rn(v.g, v.a), 255.0f)); p[2] = q8(__fmul_rn(__fmul_rn(v.b, v.a), 255.0f)); p[3] = q8(__fmul_rn(v.a, 255.0f)); }
    if (D == GB200_rgbap16){ st16(p, q16(__fmul_rn(__fmul_rn(v.r, v.a), 65535.0f))); st16(p + 2, q16(__fmul_rn(__fmul_rn(v.g, v.a), 65535.0f))); st16(p + 4, q16(__fmul_rn(__fmul_rn(v.b, v.a), 65535.0f))); st16(p + 6, q16(__fmul_rn(v.a, 65535.0f))); }
    if (D == GB200_rgbapf32){ stf(p, __fmul_rn(v.r, v.a)); stf(p + 4, __fmul_rn(v.g, v.a)); stf(p + 8, __fmul_rn(v.b, v.a)); stf(p + 12, v.a); }
}

// ---- the 8-bit block {l8,la8,rgb8,rgba8}^2 routes through rgba8 (scanline.d:160-234) ----
__device__ __forceinline__ uchar4 load_u8(int S, const uint8_t* p)
{
    switch (S) {
    case GB200_l8:   return make_uchar4(p[0], p[0], p[0], 255);
    case GB200_la8:  return make_uchar4(p[0], p[0], p[0], p[1]);
    case GB200_rgb8: return make_uchar4(p[0], p[1], p[2], 255);
    default:         return make_uchar4(p[0], p[1], p[2], p[3]);
    }
}
template <int D>
__device__ __forceinline__ void store_u8(uchar4 v, uint8_t* p)
{
    if (D == GB200_l8)   { p[0] = v.x; }
    if (D == GB200_la8)  { p[0] = v.x; p[1] = v.w; }
    if (D == GB200_rgb8) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
    if (D == GB200_rgba8){ p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w; }
}

__host__ __device__ constexpr int px_size(int t)
{
    return t == 0 ? 1 : t == 1 ? 2 : t == 2 ? 4 : t == 3 ? 2 : t == 4 ? 4 : t == 5 ? 8 : t == 6 ? 2 : t == 7 ? 4 :
           t == 8 ? 8 : t == 9 ? 3 : t == 10 ? 6 : t == 11 ? 12 : t == 12 ? 4 : t == 13 ? 8 : t == 14 ? 16 :
           t == 15 ? 4 : t == 16 ? 8 : 16;
}
__host__ __device__ constexpr bool is_8bit(int t) { return t == GB200_l8 || t == GB200_la8 || t == GB200_rgb8 || t == GB200_rgba8; }

// ---------------------------------------------------------------------------------------------
// Shared <-> global tile copies. `g` has arbitrary alignment; the shared tile is laid out so that
// (shared offset) == (g & 15), which makes the interior of the copy 16-byte aligned on both sides.
__device__ __forceinline__ void tile_g2s(uint8_t* sm, const uint8_t* g, int nbytes, int tid, int nthr)
{
    int mis = (int)((uintptr_t)g & 15);
    int head = mis ? min(16 - mis, nbytes) : 0;
    int body = (nbytes - head) >> 4;
    int tail = nbytes - head - (body << 4);
    const uint4* gv = (const uint4*)(g + head);
    uint4* sv = (uint4*)(sm + mis + head);
    for (int i = tid; i < body; i += nthr) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(gv + i));
        sv[i] = v;
    }
    if (tid < head) sm[mis + tid] = g[tid];
    if (tid < tail) sm[mis + head + (body << 4) + tid] = g[head + (body << 4) + tid];
}
__device__ __forceinline__ void tile_s2g(uint8_t* g, const uint8_t* sm, int nbytes, int tid, int nthr)
{
    int mis = (int)((uintptr_t)g & 15);
    int head = mis ? min(16 - mis, nbytes) : 0;
    int body = (nbytes - head) >> 4;
    int tail = nbytes - head - (body << 4);
    uint4* gv = (uint4*)(g + head);
    const uint4* sv = (const uint4*)(sm + mis + head);
    for (int i = tid; i < body; i += nthr) {
        uint4 v = sv[i];
        asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                     :: "l"(gv + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
    if (tid < head) g[tid] = sm[mis + tid];
    if (tid < tail) g[head + (body << 4) + tid] = sm[mis + head + (body << 4) + tid];
}

constexpr int TILE_PX = 1024;      // pixels per tile
constexpr int STAGED_THREADS = 256;

struct ConvArgs {
    const uint8_t* src; long long srcPitch;
    uint8_t* dst; long long dstPitch;
    long long width;       // pixels per row (after flattening)
    long long height;
    long long tilesPerRow;
    long long numTiles;
    int S;
};

template <int D>
__global__ void __launch_bounds__(STAGED_THREADS)
convert_staged(ConvArgs a)
{
    __shared__ __align__(16) uint8_t sm_in[TILE_PX * 16 + 32];
    __shared__ __align__(16) uint8_t sm_out[TILE_PX * px_size(D) + 32];
    const int S = a.S;
    const int ssz = px_size(S);
    constexpr int dsz = px_size(D);
    const bool via8 = is_8bit(S) && is_8bit(D);
    for (long long tile = blockIdx.x; tile < a.numTiles; tile += gridDim.x) {
        long long y = tile / a.tilesPerRow;
        long long x0 = (tile - y * a.tilesPerRow) * TILE_PX;
        int n = (int)min((long long)TILE_PX, a.width - x0);
        const uint8_t* gs = a.src + y * a.srcPitch + x0 * ssz;
        uint8_t* gd = a.dst + y * a.dstPitch + x0 * dsz;
        int smis = (int)((uintptr_t)gs & 15), dmis = (int)((uintptr_t)gd & 15);
        tile_g2s(sm_in, gs, n * ssz, threadIdx.x, STAGED_THREADS);
        __syncthreads();
        if (via8) {
            if (is_8bit(D)) {
                for (int i = threadIdx.x; i < n; i += STAGED_THREADS)
                    store_u8<D>(load_u8(S, sm_in + smis + i * ssz), sm_out + dmis + i * dsz);
            }
        } else {
            for (int i = threadIdx.x; i < n; i += STAGED_THREADS)
                store_f32<D>(load_f32(S, sm_in + smis + i * ssz), sm_out + dmis + i * dsz);
        }
        __syncthreads();
        tile_s2g(gd, sm_out, n * dsz, threadIdx.x, STAGED_THREADS);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Direct kernel: pixel sizes in {4,8,16} on both sides, pointers/pitches aligned to the pixel size.
template <int N> struct VecOf;
template <> struct VecOf<4>  { typedef uint32_t T; };
template <> struct VecOf<8>  { typedef uint2 T; };
template <> struct VecOf<16> { typedef uint4 T; };

template <int N> __device__ __forceinline__ typename VecOf<N>::T ldg_stream(const void* p);
template <> __device__ __forceinline__ uint32_t ldg_stream<4>(const void* p)
{ uint32_t v; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
template <> __device__ __forceinline__ uint2 ldg_stream<8>(const void* p)
{ uint2 v; asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p)); return v; }
template <> __device__ __forceinline__ uint4 ldg_stream<16>(const void* p)
{ uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v; }
__device__ __forceinline__ void stg_stream(void* p, uint32_t v)
{ asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void stg_stream(void* p, uint2 v)
{ asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory"); }
__device__ __forceinline__ void stg_stream(void* p, uint4 v)
{ asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }

constexpr int DIRECT_THREADS = 256;
constexpr int DIRECT_UNROLL = 8;

template <int S, int D>
__global__ void __launch_bounds__(DIRECT_THREADS)
convert_direct(const uint8_t* __restrict__ src, long long srcPitch, uint8_t* __restrict__ dst, long long dstPitch,
               long long width, long long height, long long chunksPerRow, long long numChunks)
{
    constexpr int ssz = px_size(S), dsz = px_size(D);
    typedef typename VecOf<ssz>::T SV;
    typedef typename VecOf<dsz>::T DV;
    constexpr int CHUNK = DIRECT_THREADS * DIRECT_UNROLL;
    for (long long c = blockIdx.x; c < numChunks; c += gridDim.x) {
        long long y = c / chunksPerRow;
        long long x0 = (c - y * chunksPerRow) * CHUNK + threadIdx.x;
        const uint8_t* rs = src + y * srcPitch;
        uint8_t* rd = dst + y * dstPitch;
        SV in[DIRECT_UNROLL];
#pragma unroll
        for (int k = 0; k < DIRECT_UNROLL; ++k) {
            long long x = x0 + k * DIRECT_THREADS;
            if (x < width) in[k] = ldg_stream<ssz>(rs + x * ssz);
        }
#pragma unroll
        for (int k = 0; k < DIRECT_UNROLL; ++k) {
            long long x = x0 + k * DIRECT_THREADS;
            if (x < width) {
                DV out;
                F4 v = load_f32(S, (const uint8_t*)&in[k]);
                store_f32<D>(v, (uint8_t*)&out);
                stg_stream(rd + x * dsz, out);
            }
        }
    }
}

// Same-type copy (scanline.d:37-55): rows of w*size bytes, gap bytes untouched.
__global__ void __launch_bounds__(256)
copy_rows(const uint8_t* __restrict__ src, long long srcPitch, uint8_t* __restrict__ dst, long long dstPitch,
          long long rowBytes, long long height, long long tilesPerRow, long long numTiles)
{
    __shared__ __align__(16) uint8_t sm[16384 + 32];
    for (long long tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        long long y = tile / tilesPerRow;
        long long b0 = (tile - y * tilesPerRow) * 16384;
        int n = (int)min(16384LL, rowBytes - b0);
        const uint8_t* gs = src + y * srcPitch + b0;
        uint8_t* gd = dst + y * dstPitch + b0;
        int smis = (int)((uintptr_t)gs & 15), dmis = (int)((uintptr_t)gd & 15);
        if (smis == dmis) {
            tile_g2s(sm, gs, n, threadIdx.x, 256);
            __syncthreads();
            tile_s2g(gd, sm, n, threadIdx.x, 256);
        } else {
            tile_g2s(sm, gs, n, threadIdx.x, 256);
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += 256) gd[i] = sm[smis + i];
        }
        __syncthreads();
    }
}

typedef void (*staged_fn)(ConvArgs);
template <int D> void launch_staged(ConvArgs a, int grid, cudaStream_t st) { convert_staged<D><<<grid, STAGED_THREADS, 0, st>>>(a); }
typedef void (*staged_launcher)(ConvArgs, int, cudaStream_t);
const staged_launcher g_staged[18] = {
    launch_staged<0>, launch_staged<1>, launch_staged<2>, launch_staged<3>, launch_staged<4>, launch_staged<5>,
    launch_staged<6>, launch_staged<7>, launch_staged<8>, launch_staged<9>, launch_staged<10>, launch_staged<11>,
    launch_staged<12>, launch_staged<13>, launch_staged<14>, launch_staged<15>, launch_staged<16>, launch_staged<17>};

template <int S, int D>
void launch_direct(const uint8_t* src, long long sp, uint8_t* dst, long long dp, long long w, long long h, cudaStream_t st)
{
    constexpr int CHUNK = DIRECT_THREADS * DIRECT_UNROLL;
    long long cpr = (w + CHUNK - 1) / CHUNK;
    long long nc = cpr * h;
    long long maxGrid = (long long)gb::sm_count() * 8 * 16;
    int grid = (int)(nc < maxGrid ? nc : maxGrid);
    convert_direct<S, D><<<grid, DIRECT_THREADS, 0, st>>>(src, sp, dst, dp, w, h, cpr, nc);
}

bool aligned_for(const void* p, long long pitch, int sz, long long h)
{
    if (((uintptr_t)p) % sz) return false;
    if (h > 1 && (pitch % sz)) return false;
    return true;
}

} // namespace

namespace gb {

// Device-pointer implementation shared by every entry point. Returns 1 on success, 0 on failure.
int convert_device(int srcType, const uint8_t* src, long long srcPitch, int dstType, uint8_t* dst, long long dstPitch,
                   int width, int height, cudaStream_t st)
{
    if (!ensure_device()) return 0;
    if (srcType < 0 || srcType > 17 || dstType < 0 || dstType > 17) { set_error("scanlinesConvert: unknown PixelType"); return 0; }
    if (width < 0 || height < 0) { set_error("scanlinesConvert: negative size"); return 0; }
    if (width == 0 || height == 0) return 1;
    const int ssz = px_size(srcType), dsz = px_size(dstType);
    // component alignment (the reference casts rows to ushort*/float*)
    int scomp = (srcType % 3 == 0) ? 1 : (srcType % 3 == 1) ? 2 : 4;
    int dcomp = (dstType % 3 == 0) ? 1 : (dstType % 3 == 1) ? 2 : 4;
    if (!aligned_for(src, srcPitch, scomp, height) || !aligned_for(dst, dstPitch, dcomp, height)) {
        set_error("scanlinesConvert: scanlines are not aligned to their component size");
        return 0;
    }
    long long w = width, h = height;
    // gapless on both sides: flatten to one long row
    if (srcPitch == (long long)width * ssz && dstPitch == (long long)width * dsz) { w = w * h; h = 1; }

    if (srcType == dstType) {
        long long rowBytes = w * ssz;
        long long tpr = (rowBytes + 16383) / 16384;
        long long nt = tpr * h;
        long long maxGrid = (long long)sm_count() * 16;
        int grid = (int)(nt < maxGrid ? nt : maxGrid);
        copy_rows<<<grid, 256, 0, st>>>(src, srcPitch, dst, dstPitch, rowBytes, h, tpr, nt);
        count_launch();
        GB_CUDA(cudaGetLastError());
        return 1;
    }

    bool direct_ok = aligned_for(src, srcPitch, ssz, h) && aligned_for(dst, dstPitch, dsz, h);
    if (direct_ok) {
#define GB_DIRECT(SS, DD) if (srcType == SS && dstType == DD) { launch_direct<SS, DD>(src, srcPitch, dst, dstPitch, w, h, st); count_launch(); GB_CUDA(cudaGetLastError()); return 1; }
        GB_DIRECT(GB200_rgba8, GB200_rgbaf32)
        GB_DIRECT(GB200_rgbaf32, GB200_rgba8)
        GB_DIRECT(GB200_rgba16, GB200_rgbaf32)
        GB_DIRECT(GB200_rgbaf32, GB200_rgba16)
        GB_DIRECT(GB200_rgbap8, GB200_rgbaf32)
        GB_DIRECT(GB200_rgbaf32, GB200_rgbap8)
        GB_DIRECT(GB200_la16, GB200_rgbaf32)
        GB_DIRECT(GB200_lf32, GB200_rgbaf32)
#undef GB_DIRECT
    }
    ConvArgs a;
    a.src = src; a.srcPitch = srcPitch; a.dst = dst; a.dstPitch = dstPitch;
    a.width = w; a.height = h; a.S = srcType;
    a.tilesPerRow = (w + TILE_PX - 1) / TILE_PX;
    a.numTiles = a.tilesPerRow * h;
    long long maxGrid = (long long)sm_count() * 6 * 8;
    int grid = (int)(a.numTiles < maxGrid ? a.numTiles : maxGrid);
    g_staged[dstType](a, grid, st);
    count_launch();
    GB_CUDA(cudaGetLastError());
    return 1;
}

} // namespace gb

// ---------------------------------------------------------------------------------------------
// C ABI
GB_API int gb200_pixel_type_size(int type) { return (type < 0 || type > 17) ? 0 : px_size(type); }

GB_API int gb200_scanlines_inter_type(int srcType, int dstType)
{
    return (is_8bit(srcType) && is_8bit(dstType)) ? GB200_rgba8 : GB200_rgbaf32;
}

GB_API int gb200_scanlines_convert_device(int srcType, const uint8_t* src, long long srcPitch,
                                          int dstType, uint8_t* dst, long long dstPitch,
                                          int width, int height, void* stream)
{
    gb::clear_error();
    return gb::convert_device(srcType, src, srcPitch, dstType, dst, dstPitch, width, height, (cudaStream_t)stream);
}

// Host-pointer drop-in for scanlinesConvert (scanline.d:70) / scanlinesCopy (scanline.d:37).
// interType/interBuf of the reference signature are implied (the two stages are fused on device).
GB_API int gb200_scanlines_convert(int srcType, const uint8_t* src, int srcPitch,
                                   int dstType, uint8_t* dst, int dstPitch, int width, int height)
{
    gb::clear_error();
    if (!gb::ensure_device()) return 0;
    if (srcType < 0 || srcType > 17 || dstType < 0 || dstType > 17) { gb::set_error("scanlinesConvert: unknown PixelType"); return 0; }
    if (width < 0 || height < 0) { gb::set_error("scanlinesConvert: negative size"); return 0; }
    if (width == 0 || height == 0) return 1;
    const size_t srow = (size_t)width * px_size(srcType), drow = (size_t)width * px_size(dstType);
    // device images are gapless; pad row starts to 16 bytes so that every row is vector-aligned
    const size_t sdp = (srow + 15) & ~(size_t)15, ddp = (drow + 15) & ~(size_t)15;
    gb::DevBuf ds(sdp * height), dd(ddp * height);
    if (!ds.p || !dd.p) return 0;
    // negative pitches: address the lowest row first and flip on the device side
    const uint8_t* slo = srcPitch >= 0 ? src : src + (long long)(height - 1) * srcPitch;
    uint8_t* dlo = dstPitch >= 0 ? dst : dst + (long long)(height - 1) * dstPitch;
    size_t sap = (size_t)(srcPitch >= 0 ? srcPitch : -(long long)srcPitch);
    size_t dap = (size_t)(dstPitch >= 0 ? dstPitch : -(long long)dstPitch);
    if (height == 1) { sap = srow; dap = drow; }
    const uint8_t* dsrc = ds.as<uint8_t>(); long long dsp = (long long)sdp;
    uint8_t* ddst = dd.as<uint8_t>(); long long ddpp = (long long)ddp;
    if (srcPitch < 0) { dsrc += (size_t)(height - 1) * sdp; dsp = -dsp; }
    if (dstPitch < 0) { ddst += (size_t)(height - 1) * ddp; ddpp = -ddpp; }
    // Bands of logical rows go round-robin over three streams so that the H2D copy of band k+1, the
    // kernel of band k and the D2H copy of band k-1 overlap (PCIe is full duplex; a whole-image
    // H2D -> kernel -> D2H sequence leaves one direction idle at any time).
    const size_t big = srow > drow ? srow : drow;
    int band = (int)((16u << 20) / (big ? big : 1));
    if (band < 1) band = 1;
    if (band >= height) band = height;
    // a failure after work was queued must not release ds / dd (DevBuf destructors) while copies or kernels still
    // touch them: every exit goes through the synchronisation below
#define GB_TRY(x) if (!gb::cuda_ok((x), #x, __FILE__, __LINE__)) { ok = false; break; }
    bool ok = true;
    int k = 0;
    for (int i0 = 0; i0 < height; i0 += band, ++k) {
        const int i1 = i0 + band < height ? i0 + band : height, nr = i1 - i0;
        cudaStream_t st = gb::thread_stream(k % 3);
        const size_t sm0 = srcPitch >= 0 ? (size_t)i0 : (size_t)(height - i1);   // first memory row of the band
        const size_t dm0 = dstPitch >= 0 ? (size_t)i0 : (size_t)(height - i1);
        if (sap == srow && sdp == srow) { GB_TRY(cudaMemcpyAsync(ds.as<uint8_t>() + sm0 * sdp, slo + sm0 * sap, srow * nr, cudaMemcpyHostToDevice, st)); }
        else { GB_TRY(cudaMemcpy2DAsync(ds.as<uint8_t>() + sm0 * sdp, sdp, slo + sm0 * sap, sap, srow, nr, cudaMemcpyHostToDevice, st)); }
        if (!gb::convert_device(srcType, dsrc + (long long)i0 * dsp, dsp, dstType, ddst + (long long)i0 * ddpp, ddpp, width, nr, st)) { ok = false; break; }
        if (dap == drow && ddp == drow) { GB_TRY(cudaMemcpyAsync(dlo + dm0 * dap, dd.as<uint8_t>() + dm0 * ddp, drow * nr, cudaMemcpyDeviceToHost, st)); }
        else { GB_TRY(cudaMemcpy2DAsync(dlo + dm0 * dap, dap, dd.as<uint8_t>() + dm0 * ddp, ddp, drow, nr, cudaMemcpyDeviceToHost, st)); }
    }
#undef GB_TRY
    for (int q = 0; q < 3 && q <= k; ++q) if (!gb::cuda_ok(cudaStreamSynchronize(gb::thread_stream(q)), "sync", __FILE__, __LINE__)) ok = false;
    return ok ? 1 : 0;
}
