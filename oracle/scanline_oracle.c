/*
 * scanline_oracle.c -- CPU restatement of source/gamut/scanline.d (TEST INFRASTRUCTURE ONLY).
 *
 * Every function follows the reference function of the same name; the arithmetic is written
 * in the same order, every float operation rounds to IEEE binary32 (compile with
 * -ffp-contract=off, SSE2 scalar math -- no x87, no FMA), float->integer casts truncate
 * toward zero exactly like the cvttss2si that LDC/DMD emit for cast(ubyte)/cast(ushort).
 *
 * parity: UNPINNED by the reference (no reference test asserts a converted value,
 * SURVEY.md section 8c). The oracle *is* the restated source; closed forms are cross-checked
 * in tests/test_oracle_scanline.py with numpy float32.
 */
#include "oracle.h"
#include <string.h>
#include <emmintrin.h>

/* cast(ubyte)(float) / cast(ushort)(float) as compiled for x86-64: cvttss2si then truncate. */
static inline int32_t cvtt(float v) { return _mm_cvttss_si32(_mm_set_ss(v)); }
static inline uint8_t  to_u8(float v)  { return (uint8_t)cvtt(v); }
static inline uint16_t to_u16(float v) { return (uint16_t)cvtt(v); }

/* types.d:62-86 */
int or_pixelTypeSize(int type)
{
    static const int sz[18] = {1,2,4, 2,4,8, 2,4,8, 3,6,12, 4,8,16, 4,8,16};
    if (type < 0 || type > 17) return 0;
    return sz[type];
}

/* internals/types.d:99-111 */
static int pixelTypeIs8Bit(int t) { return t == OR_l8 || t == OR_la8 || t == OR_rgb8 || t == OR_rgba8; }

/* scanline.d:25-31 (pixelTypeExpressibleInRGBA8 == pixelTypeIs8Bit, internals/types.d:144) */
int or_scanlinesInterType(int srcType, int dstType)
{
    if (pixelTypeIs8Bit(srcType) && pixelTypeIs8Bit(dstType)) return OR_rgba8;
    return OR_rgbaf32;
}

/* scanline.d:37-55 */
int or_scanlinesCopy(int type, const uint8_t* src, int srcPitch, uint8_t* dst, int dstPitch,
                     int width, int height)
{
    int scanlineBytes = or_pixelTypeSize(type) * width;
    for (int y = 0; y < height; ++y) {
        memcpy(dst, src, (size_t)scanlineBytes);
        src += srcPitch;
        dst += dstPitch;
    }
    return 1;
}

/* ---- to rgb8 helpers (scanline.d:139-154) ---- */
void or_scanline_l8_to_rgb8(const uint8_t* in, uint8_t* out, int width)
{
    for (int x = 0; x < width; ++x) { uint8_t b = in[x]; *out++ = b; *out++ = b; *out++ = b; }
}

/* ---- to rgba8 (scanline.d:160-195) ---- */
static void l8_to_rgba8(const uint8_t* in, uint8_t* out, int width)
{
    for (int x = 0; x < width; ++x) { uint8_t b = in[x]; *out++ = b; *out++ = b; *out++ = b; *out++ = 255; }
}
static void la8_to_rgba8(const uint8_t* in, uint8_t* out, int width)
{
    for (int x = 0; x < width; ++x) { uint8_t b = in[x*2]; *out++ = b; *out++ = b; *out++ = b; *out++ = in[x*2+1]; }
}
static void rgb8_to_rgba8(const uint8_t* in, uint8_t* out, int width)
{
    for (int x = 0; x < width; ++x) { *out++ = in[x*3]; *out++ = in[x*3+1]; *out++ = in[x*3+2]; *out++ = 255; }
}

/* ---- from rgba8 (scanline.d:201-234) ---- */
static void rgba8_to_l8(const uint8_t* in, uint8_t* out, int width)
{
    for (int x = 0; x < width; ++x) out[x] = in[4*x];
}
static void rgba8_to_la8(const uint8_t* in, uint8_t* out, int width)
{
    for (int x = 0; x < width; ++x) { out[2*x] = in[4*x]; out[2*x+1] = in[4*x+3]; }
}
static void rgba8_to_rgb8(const uint8_t* in, uint8_t* out, int width)
{
    for (int x = 0; x < width; ++x) { out[3*x] = in[4*x]; out[3*x+1] = in[4*x+1]; out[3*x+2] = in[4*x+2]; }
}

/* ---- to rgbaf32 (scanline.d:240-529) ----
 * T = storage type, K = normaliser (255.0f / 65535.0f / none). */
#define TO_F_L(NAME, T, NORM)                                                        \
static void NAME(const uint8_t* inScan, uint8_t* outScan, int width) {               \
    const T* s = (const T*)inScan; float* outp = (float*)outScan;                    \
    for (int x = 0; x < width; ++x) {                                                \
        float b = NORM(s[x]);                                                        \
        *outp++ = b; *outp++ = b; *outp++ = b; *outp++ = 1.0f; } }
#define TO_F_LA(NAME, T, NORM, PREMUL)                                               \
static void NAME(const uint8_t* inScan, uint8_t* outScan, int width) {               \
    const T* s = (const T*)inScan; float* outp = (float*)outScan;                    \
    for (int x = 0; x < width; ++x) {                                                \
        float b = NORM(*s); s++; float a = NORM(*s); s++;                            \
        if (PREMUL) { if (a != 0) b /= a; }                                          \
        *outp++ = b; *outp++ = b; *outp++ = b; *outp++ = a; } }
#define TO_F_RGB(NAME, T, NORM)                                                      \
static void NAME(const uint8_t* inScan, uint8_t* outScan, int width) {               \
    const T* s = (const T*)inScan; float* outp = (float*)outScan;                    \
    for (int x = 0; x < width; ++x) {                                                \
        float r = NORM(*s); s++; float g = NORM(*s); s++; float b = NORM(*s); s++;   \
        *outp++ = r; *outp++ = g; *outp++ = b; *outp++ = 1.0f; } }
#define TO_F_RGBA(NAME, T, NORM, PREMUL)                                             \
static void NAME(const uint8_t* inScan, uint8_t* outScan, int width) {               \
    const T* s = (const T*)inScan; float* outp = (float*)outScan;                    \
    for (int x = 0; x < width; ++x) {                                                \
        float r = NORM(*s); s++; float g = NORM(*s); s++;                            \
        float b = NORM(*s); s++; float a = NORM(*s); s++;                            \
        if (PREMUL) { if (a != 0) { r /= a; g /= a; b /= a; } }                      \
        *outp++ = r; *outp++ = g; *outp++ = b; *outp++ = a; } }

#define N8(v)  ((float)(v) / 255.0f)      /* ubyte / 255.0f : int->float then IEEE div */
#define N16(v) ((float)(v) / 65535.0f)
#define NF(v)  (v)

TO_F_L   (l8_to_rgbaf32,      uint8_t,  N8)       /* scanline.d:240 */
TO_F_L   (l16_to_rgbaf32,     uint16_t, N16)      /* :254 */
TO_F_L   (lf32_to_rgbaf32,    float,    NF)       /* :268 */
TO_F_LA  (la8_to_rgbaf32,     uint8_t,  N8, 0)    /* :282 */
TO_F_LA  (la16_to_rgbaf32,    uint16_t, N16, 0)   /* :297 */
TO_F_LA  (laf32_to_rgbaf32,   float,    NF, 0)    /* :312 */
TO_F_LA  (lap8_to_rgbaf32,    uint8_t,  N8, 1)    /* :328 */
TO_F_LA  (lap16_to_rgbaf32,   uint16_t, N16, 1)   /* :345 */
TO_F_LA  (lapf32_to_rgbaf32,  float,    NF, 1)    /* :362 */
TO_F_RGB (rgb8_to_rgbaf32,    uint8_t,  N8)       /* :380 */
TO_F_RGB (rgb16_to_rgbaf32,   uint16_t, N16)      /* :396 */
TO_F_RGB (rgbf32_to_rgbaf32,  float,    NF)       /* :412 */
TO_F_RGBA(rgba8_to_rgbaf32,   uint8_t,  N8, 0)    /* :428 */
TO_F_RGBA(rgba16_to_rgbaf32,  uint16_t, N16, 0)   /* :445 */
TO_F_RGBA(rgbap8_to_rgbaf32,  uint8_t,  N8, 1)    /* :462 */
TO_F_RGBA(rgbap16_to_rgbaf32, uint16_t, N16, 1)   /* :485 */
TO_F_RGBA(rgbapf32_to_rgbaf32,float,    NF, 1)    /* :508 */

/* ---- from rgbaf32 (scanline.d:539-803) ---- */
#define I(k) inp[4*x+(k)]

/* :539 / :551  -- cast(T)(0.5f + (r + g + b) * K / 3.0f), left-assoc: ((sum*K)/3) */
static void rgbaf32_to_l8(const uint8_t* inScan, uint8_t* outScan, int width)
{
    const float* inp = (const float*)inScan; uint8_t* s = outScan;
    for (int x = 0; x < width; ++x) *s++ = to_u8(0.5f + (I(0) + I(1) + I(2)) * 255.0f / 3.0f);
}
static void rgbaf32_to_l16(const uint8_t* inScan, uint8_t* outScan, int width)
{
    const float* inp = (const float*)inScan; uint16_t* s = (uint16_t*)outScan;
    for (int x = 0; x < width; ++x) *s++ = to_u16(0.5f + (I(0) + I(1) + I(2)) * 65535.0f / 3.0f);
}
/* :564 */
static void rgbaf32_to_lf32(const uint8_t* inScan, uint8_t* outScan, int width)
{
    const float* inp = (const float*)inScan; float* s = (float*)outScan;
    for (int x = 0; x < width; ++x) *s++ = (I(0) + I(1) + I(2)) / 3.0f;
}
/* :576 */
static void rgbaf32_to_la8(const uint8_t* inScan, uint8_t* outScan, int width)
{
    const float* inp = (const float*)inScan; uint8_t* s = outScan;
    for (int x = 0; x < width; ++x) {
        uint8_t b = to_u8(0.5f + (I(0) + I(1) + I(2)) * 255.0f / 3.0f);
        uint8_t a = to_u8(0.5f + I(3) * 255.0f);
        *s++ = b; *s++ = a;
    }
}
/* :593 */
static void rgbaf32_to_la16(const uint8_t* inScan, uint8_t* outScan, int width)
{
    const float* inp = (const float*)inScan; uint16_t* s = (uint16_t*)outScan;
    for (int x = 0; x < width; ++x) {
        uint16_t b = to_u16(0.5f + (I(0) + I(1) + I(2)) * 65535.0f / 3.0f);
        uint16_t a = to_u16(0.5f + I(3) * 65535.0f);
        *s++ = b; *s++ = a;
    }
}
/* :607 */
static void rgbaf32_to_laf32(const uint8_t* inScan, uint8_t* outScan, int width)
{
    const float* inp = (const float*)inScan; float* s = (float*)outScan;
    for (int x = 0; x < width; ++x) {
        float b = (I(0) + I(1) + I(2)) / 3.0f; float a = I(3);
        *s++ = b; *s++ = a;
    }
}
/* :622 -- (sum * a * 255.0f / 3.0f), left-assoc */
static void rgbaf32_to_lap8(const uint8_t* inScan, uint8_t* outScan, int width)
{
    const float* inp = (const float*)inScan; uint8_t* s = outScan;
    for (int x = 0; x < width; ++x) {
        uint8_t b = to_u8(0.5f + (I(0) + I(1) + I(2)) * I(3) * 255.0f / 3.0f);
        uint8_t a = to_u8(0.5f + I(3) * 255.0f);
        *s++ = b; *s++ = a;
    }
}
/* :639 */
static void rgbaf32_to_lap16(const uint8_t* inScan, uint8_t* outScan, int width)
{
    const float* inp = (const float*)inScan; uint16_t* s = (uint16_t*)outScan;
    for (int x = 0; x < width; ++x) {
        uint16_t b = to_u16(0.5f + (I(0) + I(1) + I(2)) * I(3) * 65535.0f / 3.0f);
        uint16_t a = to_u16(0.5f + I(3) * 65535.0f);
        *s++ = b; *s++ = a;
    }
}
/* :653 */
static void rgbaf32_to_lapf32(const uint8_t* inScan, uint8_t* outScan, int width)
{
    const float* inp = (const float*)inScan; float* s = (float*)outScan;
    for (int x = 0; x < width; ++x) {
        float b = (I(0) + I(1) + I(2)) * I(3) / 3.0f; float a = I(3);
        *s++ = b; *s++ = a;
    }
}
/* :667 */
static void rgbaf32_to_rgb8(const uint8_t* inScan, uint8_t* outScan, int width)
{
    const float* inp = (const float*)inScan; uint8_t* s = outScan;
    for (int x = 0; x < width; ++x) {
        uint8_t r = to_u8(0.5f + I(0) * 255.0f);
        uint8_t g = to_u8(0.5f + I(1) * 255.0f);
        uint8_t b = to_u8(0.5f + I(2) * 255.0f);
        *s++ = r; *s++ = g; *s++ = b;
    }
}
/* :684 */
static void rgbaf32_to_rgb16(const uint8_t* inScan, uint8_t* outScan, int width)
{
    const float* inp = (const float*)inScan; uint16_t* s = (uint16_t*)outScan;
    for (int x = 0; x < width; ++x) {
        uint16_t r = to_u16(0.5f + I(0) * 65535.0f);
        uint16_t g = to_u16(0.5f + I(1) * 65535.0f);
        uint16_t b = to_u16(0.5f + I(2) * 65535.0f);
        *s++ = r; *s++ = g; *s++ = b;
    }
}
/* :700 */
static void rgbaf32_to_rgbf32(const uint8_t* inScan, uint8_t* outScan, int width)
{
    const float* inp = (const float*)inScan; float* s = (float*)outScan;
    for (int x = 0; x < width; ++x) { *s++ = I(0); *s++ = I(1); *s++ = I(2); }
}
/* :713 */
static void rgbaf32_to_rgba8(const uint8_t* inScan, uint8_t* outScan, int width)
{
    const float* inp = (const float*)inScan; uint8_t* s = outScan;
    for (int x = 0; x < width; ++x) {
        uint8_t r = to_u8(0.5f + I(0) * 255.0f);
        uint8_t g = to_u8(0.5f + I(1) * 255.0f);
        uint8_t b = to_u8(0.5f + I(2) * 255.0f);
        uint8_t a = to_u8(0.5f + I(3) * 255.0f);
        *s++ = r; *s++ = g; *s++ = b; *s++ = a;
    }
}
/* :731 */
static void rgbaf32_to_rgba16(const uint8_t* inScan, uint8_t* outScan, int width)
{
    const float* inp = (const float*)inScan; uint16_t* s = (uint16_t*)outScan;
    for (int x = 0; x < width; ++x) {
        uint16_t r = to_u16(0.5f + I(0) * 65535.0f);
        uint16_t g = to_u16(0.5f + I(1) * 65535.0f);
        uint16_t b = to_u16(0.5f + I(2) * 65535.0f);
        uint16_t a = to_u16(0.5f + I(3) * 65535.0f);
        *s++ = r; *s++ = g; *s++ = b; *s++ = a;
    }
}
/* :755 -- 0.5f + c * a * 255.0f, left-assoc: ((c*a)*255) */
static void rgbaf32_to_rgbap8(const uint8_t* inScan, uint8_t* outScan, int width)
{
    const float* inp = (const float*)inScan; uint8_t* s = outScan;
    for (int x = 0; x < width; ++x) {
        uint8_t r = to_u8(0.5f + I(0) * I(3) * 255.0f);
        uint8_t g = to_u8(0.5f + I(1) * I(3) * 255.0f);
        uint8_t b = to_u8(0.5f + I(2) * I(3) * 255.0f);
        uint8_t a = to_u8(0.5f + I(3) * 255.0f);
        *s++ = r; *s++ = g; *s++ = b; *s++ = a;
    }
}
/* :773 */
static void rgbaf32_to_rgbap16(const uint8_t* inScan, uint8_t* outScan, int width)
{
    const float* inp = (const float*)inScan; uint16_t* s = (uint16_t*)outScan;
    for (int x = 0; x < width; ++x) {
        uint16_t r = to_u16(0.5f + I(0) * I(3) * 65535.0f);
        uint16_t g = to_u16(0.5f + I(1) * I(3) * 65535.0f);
        uint16_t b = to_u16(0.5f + I(2) * I(3) * 65535.0f);
        uint16_t a = to_u16(0.5f + I(3) * 65535.0f);
        *s++ = r; *s++ = g; *s++ = b; *s++ = a;
    }
}
/* :791 */
static void rgbaf32_to_rgbapf32(const uint8_t* inScan, uint8_t* outScan, int width)
{
    const float* inp = (const float*)inScan; float* s = (float*)outScan;
    for (int x = 0; x < width; ++x) {
        float a = I(3);
        *s++ = I(0) * a; *s++ = I(1) * a; *s++ = I(2) * a; *s++ = a;
    }
}
#undef I

/* ---- BMP ordering (scanline.d:812-836) ---- */
void or_scanline_rgba8_to_bgra8(const uint8_t* in, uint8_t* out, int width)
{
    for (int x = 0; x < width; ++x) {
        out[4*x+0] = in[4*x+2]; out[4*x+1] = in[4*x+1]; out[4*x+2] = in[4*x+0]; out[4*x+3] = in[4*x+3];
    }
}
void or_scanline_rgb8_to_bgr8(const uint8_t* in, uint8_t* out, int width)
{
    for (int x = 0; x < width; ++x) {
        out[3*x+0] = in[3*x+2]; out[3*x+1] = in[3*x+1]; out[3*x+2] = in[3*x+0];
    }
}

/* scanline.d:841-885 */
static int convertToIntermediateScanline(int srcType, const uint8_t* src, int dstType, uint8_t* dest, int width)
{
    if (dstType == OR_rgba8) {
        switch (srcType) {
        case OR_l8:    l8_to_rgba8(src, dest, width); break;
        case OR_la8:   la8_to_rgba8(src, dest, width); break;
        case OR_rgb8:  rgb8_to_rgba8(src, dest, width); break;
        case OR_rgba8: memcpy(dest, src, (size_t)width * 4); break;
        default: return 0; /* assert(false) in the reference */
        }
    } else if (dstType == OR_rgbaf32) {
        switch (srcType) {
        case OR_l8:      l8_to_rgbaf32(src, dest, width); break;
        case OR_l16:     l16_to_rgbaf32(src, dest, width); break;
        case OR_lf32:    lf32_to_rgbaf32(src, dest, width); break;
        case OR_la8:     la8_to_rgbaf32(src, dest, width); break;
        case OR_la16:    la16_to_rgbaf32(src, dest, width); break;
        case OR_laf32:   laf32_to_rgbaf32(src, dest, width); break;
        case OR_lap8:    lap8_to_rgbaf32(src, dest, width); break;
        case OR_lap16:   lap16_to_rgbaf32(src, dest, width); break;
        case OR_lapf32:  lapf32_to_rgbaf32(src, dest, width); break;
        case OR_rgb8:    rgb8_to_rgbaf32(src, dest, width); break;
        case OR_rgb16:   rgb16_to_rgbaf32(src, dest, width); break;
        case OR_rgbf32:  rgbf32_to_rgbaf32(src, dest, width); break;
        case OR_rgba8:   rgba8_to_rgbaf32(src, dest, width); break;
        case OR_rgba16:  rgba16_to_rgbaf32(src, dest, width); break;
        case OR_rgbaf32: memcpy(dest, src, (size_t)width * 16); break;
        case OR_rgbap8:  rgbap8_to_rgbaf32(src, dest, width); break;
        case OR_rgbap16: rgbap16_to_rgbaf32(src, dest, width); break;
        case OR_rgbapf32:rgbapf32_to_rgbaf32(src, dest, width); break;
        default: return 0;
        }
    } else return 0;
    return 1;
}

/* scanline.d:887-930 */
static int convertFromIntermediate(int srcType, const uint8_t* src, int dstType, uint8_t* dest, int width)
{
    if (srcType == OR_rgba8) {
        switch (dstType) {
        case OR_l8:    rgba8_to_l8(src, dest, width); break;
        case OR_la8:   rgba8_to_la8(src, dest, width); break;
        case OR_rgb8:  rgba8_to_rgb8(src, dest, width); break;
        case OR_rgba8: memcpy(dest, src, (size_t)width * 4); break;
        default: return 0;
        }
    } else if (srcType == OR_rgbaf32) {
        switch (dstType) {
        case OR_l8:      rgbaf32_to_l8(src, dest, width); break;
        case OR_l16:     rgbaf32_to_l16(src, dest, width); break;
        case OR_lf32:    rgbaf32_to_lf32(src, dest, width); break;
        case OR_la8:     rgbaf32_to_la8(src, dest, width); break;
        case OR_la16:    rgbaf32_to_la16(src, dest, width); break;
        case OR_laf32:   rgbaf32_to_laf32(src, dest, width); break;
        case OR_lap8:    rgbaf32_to_lap8(src, dest, width); break;
        case OR_lap16:   rgbaf32_to_lap16(src, dest, width); break;
        case OR_lapf32:  rgbaf32_to_lapf32(src, dest, width); break;
        case OR_rgb8:    rgbaf32_to_rgb8(src, dest, width); break;
        case OR_rgb16:   rgbaf32_to_rgb16(src, dest, width); break;
        case OR_rgbf32:  rgbaf32_to_rgbf32(src, dest, width); break;
        case OR_rgba8:   rgbaf32_to_rgba8(src, dest, width); break;
        case OR_rgba16:  rgbaf32_to_rgba16(src, dest, width); break;
        case OR_rgbaf32: memcpy(dest, src, (size_t)width * 16); break;
        case OR_rgbap8:  rgbaf32_to_rgbap8(src, dest, width); break;
        case OR_rgbap16: rgbaf32_to_rgbap16(src, dest, width); break;
        case OR_rgbapf32:rgbaf32_to_rgbapf32(src, dest, width); break;
        default: return 0;
        }
    } else return 0;
    return 1;
}

/* scanline.d:70-121 */
int or_scanlinesConvert(int srcType, const uint8_t* src, int srcPitch,
                        int dstType, uint8_t* dst, int dstPitch,
                        int width, int height, int interType, uint8_t* interBuf)
{
    if (srcType == dstType)
        return or_scanlinesCopy(srcType, src, srcPitch, dst, dstPitch, width, height);
    if (srcType < 0 || srcType > 17 || dstType < 0 || dstType > 17) return 0;

    if (srcType == interType) {
        for (int y = 0; y < height; ++y) {
            if (!convertFromIntermediate(srcType, src, dstType, dst, width)) return 0;
            src += srcPitch; dst += dstPitch;
        }
    } else if (dstType == interType) {
        for (int y = 0; y < height; ++y) {
            if (!convertToIntermediateScanline(srcType, src, dstType, dst, width)) return 0;
            src += srcPitch; dst += dstPitch;
        }
    } else {
        for (int y = 0; y < height; ++y) {
            if (!convertToIntermediateScanline(srcType, src, interType, interBuf, width)) return 0;
            if (!convertFromIntermediate(interType, interBuf, dstType, dst, width)) return 0;
            src += srcPitch; dst += dstPitch;
        }
    }
    return 1;
}

void or_free(void* p) { extern void free(void*); free(p); }
