/*
 * oracle.h -- CPU restatement of the AuburnSounds/gamut hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This directory is the parity oracle: plain C that follows the reference's D source
 * function by function (each function cites reference file:line). It is used only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * The product (gamut_b200/) never links, imports or calls anything in here.
 *
 * The reference is D; no D compiler exists in the build image, so the reference itself
 * cannot be compiled (oracle/_ref is therefore absent). The oracle is pinned against the
 * reference's own fixtures and KATs (tests/golden/, see tests/test_oracle_*.py).
 */
#ifndef GAMUT_ORACLE_H
#define GAMUT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* PixelType values: source/gamut/types.d:32-59 */
enum {
    OR_unknown = -1,
    OR_l8 = 0, OR_l16, OR_lf32,
    OR_la8, OR_la16, OR_laf32,
    OR_lap8, OR_lap16, OR_lapf32,
    OR_rgb8, OR_rgb16, OR_rgbf32,
    OR_rgba8, OR_rgba16, OR_rgbaf32,
    OR_rgbap8, OR_rgbap16, OR_rgbapf32
};

/* ---- scanline.d ---- */
int  or_pixelTypeSize(int type);                       /* types.d:62-86 */
int  or_scanlinesInterType(int srcType, int dstType);  /* scanline.d:25-31 */
int  or_scanlinesCopy(int type, const uint8_t* src, int srcPitch, uint8_t* dst, int dstPitch,
                      int width, int height);          /* scanline.d:37-55 */
int  or_scanlinesConvert(int srcType, const uint8_t* src, int srcPitch,
                         int dstType, uint8_t* dst, int dstPitch,
                         int width, int height, int interType, uint8_t* interBuf); /* scanline.d:70-121 */
void or_scanline_rgba8_to_bgra8(const uint8_t* in, uint8_t* out, int width); /* scanline.d:812 */
void or_scanline_rgb8_to_bgr8(const uint8_t* in, uint8_t* out, int width);   /* scanline.d:826 */
void or_scanline_l8_to_rgb8(const uint8_t* in, uint8_t* out, int width);     /* scanline.d:139 */

/* ---- PNG (stbdec.d) ---- */
typedef struct {
    int width, height;
    int channels;       /* channels in the returned buffer */
    int file_channels;  /* img_n reported by the file (before req_comp) */
    int bits;           /* 8 or 16: bits per channel of the returned buffer */
    float ppmX, ppmY, pixelRatio;
} or_png_info;

/* stbi_load_from_callbacks / stbi_load_16_from_callbacks (stbdec.d:713-735) over a memory
 * buffer. want16: 0 => 8-bit result, 1 => 16-bit result. Returns malloc'd pixels or NULL. */
uint8_t* or_png_load(const uint8_t* data, size_t len, int req_comp, int want16, or_png_info* info);
int      or_png_is16(const uint8_t* data, size_t len);   /* stbi__png_is16 stbdec.d:2104 */
/* stbi__create_png_image_raw (stbdec.d:1406-1635), depth 8/16, non-interlaced, exposed for
 * unfilter-only parity: raw = inflated stream. Returns 1 on success. */
int or_png_unfilter(const uint8_t* raw, size_t raw_len, int img_n, int out_n, int w, int h,
                    int depth, uint8_t* out);
/* inflate wrapper semantic of stbdec.d:1267-1321 (zlib is the inflate engine; miniz is
 * absent from the reference tree). Returns malloc'd buffer, sets *outlen. */
uint8_t* or_zlib_decode(const uint8_t* in, size_t inlen, size_t guess, int parse_header, size_t* outlen);

/* ---- TGA (codecs/tga.d:313-646 as plugins/tga.d:45-105 calls it) ---- */
uint8_t* or_tga_load(const uint8_t* data, size_t len, int* width, int* height, int* comp);
uint8_t* or_tga_encode(const uint8_t* data, int type, int width, int height, int pitchBytes, int* out_len);   /* plugins/tga.d:123, codecs/tga.d:62-292 */

/* ---- BMP (stbdec.d:2112-2510) and format detection (image.d:1045-1061, plugins' detect procs) ---- */
uint8_t* or_bmp_encode(const uint8_t* data, int type, int width, int height, int pitchBytes, float ppmX, float ppmY, int* out_len);   /* plugins/bmp.d:166, codecs/bmpenc.d:25-113 */
uint8_t* or_bmp_load(const uint8_t* data, size_t len, int req_comp, int* x, int* y, int* comp,
                     float* ppmX, float* ppmY, float* pixelRatio);
int or_identify_format(const uint8_t* data, size_t len);      /* ImageFormat value (types.d:14-28) or -1 */

/* ---- JPEG (jpegload.d) ---- */
uint8_t* or_jpeg_load(const uint8_t* data, size_t len, int req_comps, int* w, int* h,
                      int* actual_comps, float* par, float* dpiY);

/* ---- QOI (qoi.d) ---- */
typedef struct { uint32_t width, height; uint8_t channels, colorspace; } or_qoi_desc;
uint8_t* or_qoi_decode(const uint8_t* data, int size, or_qoi_desc* desc, int channels);  /* qoi.d:448 */
uint8_t* or_qoi_encode(const uint8_t* data, uint32_t width, uint32_t height, int pitchBytes, int channels, int colorspace, int* out_len);    /* qoi.d:295 */

/* ---- QOIX family (qoi2avg.d, qoiplane.d, qoiplane10.d, qoi10b.d, plugins/qoix.d, lz4.d) ---- */
typedef struct {
    uint32_t width, height;
    int32_t  pitchBytes;
    uint8_t  channels, bitdepth, colorspace, compression;
    float    pixelAspectRatio, resolutionY;
} or_qoix_desc;  /* qoi2avg.d:276-287 */

uint8_t* or_qoix_lz4_decode(const uint8_t* data, int size, or_qoix_desc* desc, int flags, int* decodedType); /* plugins/qoix.d:350 */
uint8_t* or_qoix_lz4_encode(const uint8_t* pixels, const or_qoix_desc* desc, int force_lz4, int* out_len);   /* plugins/qoix.d:251 */
uint8_t* or_qoiplane10_decode(const uint8_t* data, int size, or_qoix_desc* desc, int channels); /* qoiplane10.d:317 */
uint8_t* or_qoiplane10_encode(const uint8_t* pixels, const or_qoix_desc* desc, int* out_len);   /* qoiplane10.d:99 */
uint8_t* or_qoiplane_decode(const uint8_t* data, int size, or_qoix_desc* desc, int channels);   /* qoiplane.d:377 */
uint8_t* or_qoiplane_encode(const uint8_t* pixels, const or_qoix_desc* desc, int* out_len);     /* qoiplane.d:99 */
uint8_t* or_qoix_decode(const uint8_t* data, int size, or_qoix_desc* desc, int channels);       /* qoi2avg.d:625 */
uint8_t* or_qoix_encode(const uint8_t* pixels, const or_qoix_desc* desc, int* out_len);         /* qoi2avg.d:376 */
uint8_t* or_qoi10b_decode(const uint8_t* data, int size, or_qoix_desc* desc, int channels);     /* qoi10b.d:504 */
uint8_t* or_qoi10b_encode(const uint8_t* pixels, const or_qoix_desc* desc, int* out_len);       /* qoi10b.d:136 */
int or_lz4_decompress_fast(const uint8_t* src, uint8_t* dst, int originalSize);  /* lz4.d:976 */
int or_lz4_compress(const uint8_t* src, uint8_t* dst, int srcSize);              /* lz4.d:544 */
int or_lz4_compress_bound(int isize);                                            /* lz4.d:68 */

/* test hooks: single arithmetic kernels, compared with vectors generated from the reference's source text
 * (tests/golden/gen_from_reference.py, tests/test_oracle_reference_text.py) */
void or_test_jpeg_idct(const int16_t* src, int max_zag, uint8_t* dst);           /* idct, jpegload.d:308-376 */
void or_test_jpeg_upsample(const int16_t* src, uint8_t* dst256);                 /* transform_mcu_expand chroma, :2150-2251 */
void or_test_jpeg_ycc(int y, int cb, int cr, uint8_t* rgb, int* tables);         /* create_look_ups + H1V1Convert pixel */
int  or_test_loco_predict10(int left, int top, int topleft);                     /* locoPredict, qoiplane10.d:84-96 */
int  or_test_loco8(int a, int b, int c);                                         /* qoi2avg.d:863-897, one lane */
int  or_test_loco10(int a, int b, int c);                                        /* qoi10b.d:871-903, one lane */

void or_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
