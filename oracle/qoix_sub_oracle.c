/*
 * qoix_sub_oracle.c -- CPU restatement of the remaining QOIX sub-decoders and of two of their encoders (TEST
 * INFRASTRUCTURE ONLY; see oracle.h):
 *
 *   or_qoix_decode      qoix_decode      source/gamut/codecs/qoi2avg.d:625-839  (8-bit RGB/RGBA, "QOI2AVG")
 *                       locoIntraPredictionSIMD  qoi2avg.d:863-897
 *   or_qoiplane_decode  qoiplane_decode  source/gamut/codecs/qoiplane.d:377-541 (8-bit L/LA)
 *   or_qoi10b_decode    qoi10b_decode    source/gamut/codecs/qoi10b.d:504-869   (10-bit, 1-4 channels)
 *                       locoIntraPredictionSIMD  qoi10b.d:871-903
 *
 * Restatement choices (equivalent on valid streams; the reference trusts its input):
 *  - reads past the end of the input yield 0xFF bytes (the reference reads out of bounds). 0xFF is END in
 *    QOI2AVG / QOI-10b and "repeat until the end" in QOI-Plane;
 *  - memory the reference leaves uninitialised (pixels after END, the scanline double buffer) is zero.
 *
 * parity: the decoders are pinned by (a) independent minimal encoders written from the format description
 * (tests/qoixsynth.py: image -> stream -> or_*_decode == image), (b) structural review against the cited lines and,
 * since round 2, (c) the reference's round-trip property through the restated encoders of two of the three codecs
 * (or_qoiplane_encode, qoiplane.d:109-375; or_qoix_encode, qoi2avg.d:376-617, at the end of this file; qoi10b_encode is
 * not restated). No reference-made vectors exist for these streams.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

#define QOIX_MAGIC 0x716F6978u
#define QOIX_HEADER_SIZE 25
#define QOIX_PIXELS_MAX 400000000u

typedef struct { const uint8_t* b; int size; } src_t;
static inline int rdb(const src_t* s, int p) { return (p >= 0 && p < s->size) ? s->b[p] : 0xFF; }
static uint32_t be32(const uint8_t* b, int p) { return ((uint32_t)b[p] << 24) | ((uint32_t)b[p + 1] << 16) | ((uint32_t)b[p + 2] << 8) | b[p + 3]; }
static float be32f(const uint8_t* b, int p) { uint32_t v = be32(b, p); float f; memcpy(&f, &v, 4); return f; }

static int read_header(const uint8_t* bytes, or_qoix_desc* desc, int* version)
{
    desc->width = be32(bytes, 4); desc->height = be32(bytes, 8);
    *version = bytes[12]; desc->channels = bytes[13]; desc->bitdepth = bytes[14]; desc->colorspace = bytes[15];
    desc->compression = bytes[16];
    desc->pixelAspectRatio = be32f(bytes, 17); desc->resolutionY = be32f(bytes, 21);
    return be32(bytes, 0) == QOIX_MAGIC;
}

/* ------------------------------------------------------------------------------------------ */
/* QOI2AVG, qoi2avg.d                                                                            */
typedef struct { uint8_t r, g, b, a; } rgba8;

static int clamp255(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }
/* qoi2avg.d:863-897: P = A+B-C; where C >= max(A,B) take min; then where C <= min(A,B) take max (applied
   second, so it wins when both hold); saturate to 0..255 (_mm_packus_epi16) */
static int loco8(int a, int b, int c)
{
    int mx = a > b ? a : b, mn = a < b ? a : b;
    int p = a + b - c;
    if (c >= mx) p = mn;
    if (c <= mn) p = mx;
    return clamp255(p);
}

int or_test_loco8(int a, int b, int c) { return loco8(a, b, c); }          /* test hook */

uint8_t* or_qoix_decode(const uint8_t* data, int size, or_qoix_desc* desc, int channels)   /* qoi2avg.d:625 */
{
    if (!data || !desc || (channels != 0 && channels != 3 && channels != 4) || size < QOIX_HEADER_SIZE + 4) return NULL;
    int version;
    int magic_ok = read_header(data, desc, &version);
    if (desc->width == 0 || desc->height == 0 || desc->channels < 3 || desc->channels > 4 || desc->colorspace > 2 ||
        desc->bitdepth != 8 || version > 1 || desc->compression != 0 || !magic_ok ||
        desc->height >= QOIX_PIXELS_MAX / desc->width) return NULL;
    if (channels == 0) channels = desc->channels;
    const int W = (int)desc->width, H = (int)desc->height;
    desc->pitchBytes = W * channels;
    uint8_t* pixels = (uint8_t*)calloc((size_t)W * H * channels + 1, 1);
    rgba8* cur = (rgba8*)calloc((size_t)W, 4), *last = (rgba8*)calloc((size_t)W, 4);
    if (!pixels || !cur || !last) { free(pixels); free(cur); free(last); return NULL; }
    src_t S = {data, size};
    rgba8 index[64]; memset(index, 0, sizeof(index));
    rgba8 px = {0, 0, 0, 255}, ref;
    int p = QOIX_HEADER_SIZE, run = 0, index_pos = 0;
    const int chunks_len = size - 4;
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            if (run > 0) run--;
            else if (p < chunks_len) {
                ref = px;
                if (y > 0) {
                    if (x == 0) { ref.r = last[0].r; ref.g = last[0].g; ref.b = last[0].b; }
                    else {
                        ref.r = (uint8_t)loco8(px.r, last[x].r, last[x - 1].r);
                        ref.g = (uint8_t)loco8(px.g, last[x].g, last[x - 1].g);
                        ref.b = (uint8_t)loco8(px.b, last[x].b, last[x - 1].b);
                    }
                }
                int end = 0;
                for (;;) {
                    int b1 = rdb(&S, p++);
                    if (b1 < 0x80) {                                      /* LUMA */
                        int vg = ((b1 >> 4) & 7) - 4;
                        px.g = (uint8_t)(ref.g + vg);
                        int bias = vg < 0 ? 1 : 2;
                        px.r = (uint8_t)(ref.r + vg - bias + ((b1 >> 2) & 3));
                        px.b = (uint8_t)(ref.b + vg - bias + (b1 & 3));
                        index[index_pos++ & 63] = px;
                    } else if (b1 < 0xc0) px = index[b1 & 63];            /* INDEX */
                    else if (b1 < 0xe0) {                                 /* LUMA2 */
                        int b2 = rdb(&S, p++);
                        int vg = (b1 & 0x1f) - 16;
                        px.r = (uint8_t)(ref.r + vg - 8 + ((b2 >> 4) & 0x0f));
                        px.g = (uint8_t)(ref.g + vg);
                        px.b = (uint8_t)(ref.b + vg - 8 + (b2 & 0x0f));
                        index[index_pos++ & 63] = px;
                    } else if (b1 < 0xe8) {                               /* LUMA3 */
                        int dv = (b1 << 8) | rdb(&S, p++);
                        dv = (dv << 8) | rdb(&S, p++);
                        int vg = ((dv >> 12) & 0x7f) - 64;
                        px.r = (uint8_t)(ref.r + vg + ((dv >> 6) & 0x3f) - 32);
                        px.g = (uint8_t)(ref.g + vg);
                        px.b = (uint8_t)(ref.b + vg + (dv & 0x3f) - 32);
                        index[index_pos++ & 63] = px;
                    } else if (b1 < 0xf0) { px.a = (uint8_t)(px.a + (b1 & 7) - 4); continue; }   /* ADIFF */
                    else if (b1 < 0xf8) run = b1 & 7;                     /* RUN */
                    else if (b1 < 0xfc) run = ((b1 & 3) << 8) | rdb(&S, p++);   /* RUN2 */
                    else if (b1 == 0xfc) { int v = rdb(&S, p++); px.r = px.g = px.b = (uint8_t)v; index[index_pos++ & 63] = px; }
                    else if (b1 == 0xfd) { px.r = (uint8_t)rdb(&S, p++); px.g = (uint8_t)rdb(&S, p++); px.b = (uint8_t)rdb(&S, p++); index[index_pos++ & 63] = px; }
                    else if (b1 == 0xfe) { px.r = (uint8_t)rdb(&S, p++); px.g = (uint8_t)rdb(&S, p++); px.b = (uint8_t)rdb(&S, p++); px.a = (uint8_t)rdb(&S, p++); index[index_pos++ & 63] = px; }
                    else end = 1;                                         /* END: leaves the x loop only */
                    break;
                }
                if (end) break;
            }
            cur[x] = px;
        }
        uint8_t* line = pixels + (size_t)desc->pitchBytes * y;
        if (channels == 4) memcpy(line, cur, (size_t)W * 4);
        else for (int x = 0; x < W; ++x) { line[x * 3] = cur[x].r; line[x * 3 + 1] = cur[x].g; line[x * 3 + 2] = cur[x].b; }
        rgba8* t = cur; cur = last; last = t;
    }
    free(cur); free(last);
    return pixels;
}

/* ------------------------------------------------------------------------------------------ */
/* QOI-Plane (8-bit L / LA), qoiplane.d                                                          */
typedef struct { src_t S; int p; int hi; } nibr;
static int readNibble(nibr* r)                                   /* qoiplane.d:428-438 */
{
    int v;
    if (r->hi) v = rdb(&r->S, r->p) >> 4; else v = rdb(&r->S, r->p++) & 0xf;
    r->hi = !r->hi;
    return v;
}
static int readUbyte(nibr* r) { int hi = readNibble(r) << 4; return hi | readNibble(r); }

uint8_t* or_qoiplane_decode(const uint8_t* data, int size, or_qoix_desc* desc, int channels)   /* qoiplane.d:377 */
{
    if (size < QOIX_HEADER_SIZE + 4) return NULL;      /* (channels < 0 && channels > 2) is never true in the reference */
    int version;
    int magic_ok = read_header(data, desc, &version);
    if (desc->width == 0 || desc->height == 0 || desc->channels < 1 || desc->channels > 2 || desc->colorspace > 1 ||
        desc->bitdepth != 8 || version > 1 || desc->compression != 0 || !magic_ok ||
        desc->height >= QOIX_PIXELS_MAX / desc->width) return NULL;
    if (channels == 0) channels = desc->channels;
    const int W = (int)desc->width, H = (int)desc->height;
    desc->pitchBytes = W * channels;
    const int num_pixels = W * H;
    uint8_t* pixels = (uint8_t*)calloc((size_t)num_pixels * channels + 1, 1);
    if (!pixels) return NULL;
    nibr R = {{data, size}, QOIX_HEADER_SIZE, 1};
    int l = 0, a = 255, decoded = 0, run = 0;
    for (int y = 0; y < H; ++y) {
        uint8_t* line = pixels + (size_t)desc->pitchBytes * y;
        const uint8_t* above = y > 0 ? pixels + (size_t)desc->pitchBytes * (y - 1) : NULL;
        for (int x = 0; x < W; ++x) {
            const int ref_l = l, ref_a = a;
            if (run > 0) run--;
            else if (decoded < num_pixels) {
                for (;;) {
                    int op = readNibble(&R);
                    if ((op & 0xf) == 0xf) { run = readUbyte(&R) + 3; if (run == 258) run = 0x7fffffff; }
                    else if ((op & 0xc) == 0xc) run = op & 3;
                    else {
                        int top = y > 0 ? above[x * channels] : ref_l;
                        int avg = (top + ref_l + 1) / 2;
                        if ((op & 0x8) == 0) l = (uint8_t)(avg + op - 4);
                        else if ((op & 0xe) == 0x8) { int v = ((op & 1) << 4) + readNibble(&R); l = (uint8_t)(avg + v - 16); }
                        else if (op == 0xa) l = readUbyte(&R);
                        else {      /* 0xb */
                            int diff = readNibble(&R);
                            if (diff == 0) { l = readUbyte(&R); a = readUbyte(&R); }
                            else { a = (uint8_t)(ref_a + diff - 8); continue; }
                        }
                    }
                    break;
                }
                decoded++;
            }
            if (channels == 1) line[x] = (uint8_t)l;
            else { line[x * 2] = (uint8_t)l; line[x * 2 + 1] = (uint8_t)a; }
        }
    }
    return pixels;
}

/* ------------------------------------------------------------------------------------------ */
/* QOI-10b, qoi10b.d                                                                              */
typedef struct { uint16_t r, g, b, a; } rgba10;
typedef struct { src_t S; int p; int currentBit; } bitr10;
static int read2(bitr10* r)                                       /* qoi10b.d:581-594 */
{
    int bit = (rdb(&r->S, r->p) >> (r->currentBit - 1)) & 3;
    r->currentBit -= 2;
    if (r->currentBit == -1) { r->currentBit = 7; r->p++; }
    return bit;
}
static uint32_t readBits10(bitr10* r, int n) { uint32_t v = 0; for (int b = 0; b < n; b += 2) v = (v << 2) | (uint32_t)read2(r); return v; }
static void rewind1(bitr10* r) { if (r->currentBit == 7) { r->p--; r->currentBit = -1; } r->currentBit++; }
static int sx(uint32_t v, int bits) { return (int)(v << (32 - bits)) >> (32 - bits); }
static int clamp1023(int v) { return v < 0 ? 0 : v > 1023 ? 1023 : v; }
static int loco10(int a, int b, int c)                            /* qoi10b.d:871-903 (16-bit lanes: values are 0..1023) */
{
    int mx = a > b ? a : b, mn = a < b ? a : b;
    int p = a + b - c;
    if (c >= mx) p = mn;
    if (c <= mn) p = mx;
    return clamp1023(p);
}
int or_test_loco10(int a, int b, int c) { return loco10(a, b, c); }        /* test hook */

uint8_t* or_qoi10b_decode(const uint8_t* data, int size, or_qoix_desc* desc, int channels)   /* qoi10b.d:504 */
{
    if (!data || !desc || channels < 0 || channels > 4 || size < QOIX_HEADER_SIZE + 5) return NULL;
    int version;
    int magic_ok = read_header(data, desc, &version);
    if (desc->width == 0 || desc->height == 0 || desc->channels < 1 || desc->channels > 4 || desc->colorspace > 2 ||
        desc->bitdepth != 10 || version > 2 || desc->compression != 0 || !magic_ok ||
        desc->height >= QOIX_PIXELS_MAX / desc->width) return NULL;
    const int streamChannels = desc->channels;
    if (channels == 0) channels = streamChannels;
    const int W = (int)desc->width, H = (int)desc->height;
    desc->pitchBytes = W * channels * 2;
    uint8_t* pixels = (uint8_t*)calloc((size_t)desc->pitchBytes * H + 2, 1);
    rgba10* cur = (rgba10*)calloc((size_t)W, sizeof(rgba10)), *last = (rgba10*)calloc((size_t)W, sizeof(rgba10));
    if (!pixels || !cur || !last) { free(pixels); free(cur); free(last); return NULL; }
    bitr10 R = {{data, size}, QOIX_HEADER_SIZE, 7};
    const int grey = streamChannels == 1 || streamChannels == 2;
    rgba10 px = {0, 0, 0, 1023}, ref;
    int run = 0, finished = 0;
    for (int y = 0; y < H && !finished; ++y) {
        for (int x = 0; x < W; ++x) {
            ref = px;
            if (run > 0) run--;
            else {
                if (y > 0) {
                    if (version >= 2) {
                        if (x == 0) { ref.r = last[0].r; ref.g = last[0].g; ref.b = last[0].b; }
                        else {
                            rgba10 a = ref;
                            ref.r = (uint16_t)loco10(a.r, last[x].r, last[x - 1].r);
                            ref.g = (uint16_t)loco10(a.g, last[x].g, last[x - 1].g);
                            ref.b = (uint16_t)loco10(a.b, last[x].b, last[x - 1].b);
                        }
                    } else {
                        ref.r = (uint16_t)((ref.r + last[x].r + 1) >> 1);
                        ref.g = (uint16_t)((ref.g + last[x].g + 1) >> 1);
                        ref.b = (uint16_t)((ref.b + last[x].b + 1) >> 1);
                    }
                }
                for (;;) {
                    int op = (int)readBits10(&R, 8);
                    if (op < 0x80) {                                       /* LUMA */
                        int vg = sx((uint32_t)(op >> 2) & 31, 5);
                        px.g = (uint16_t)((ref.g + vg) & 1023);
                        if (!grey) {
                            int vg_r = sx((uint32_t)((op & 3) << 2) | readBits10(&R, 2), 4);
                            int vg_b = sx(readBits10(&R, 4), 4);
                            px.r = (uint16_t)((ref.r + vg + vg_r) & 1023); px.b = (uint16_t)((ref.b + vg + vg_b) & 1023);
                        } else { rewind1(&R); rewind1(&R); px.r = px.g; px.b = px.g; }
                    } else if (op < 0xc0) {                                /* LUMA0 */
                        int vg = sx((uint32_t)(op >> 2) & 15, 4);
                        px.g = (uint16_t)((ref.g + vg) & 1023);
                        if (!grey) {
                            uint32_t remain = readBits10(&R, 4);
                            int vg_r = sx((uint32_t)((op & 3) << 1) | (remain >> 3), 3);
                            int vg_b = sx(remain & 7, 3);
                            px.r = (uint16_t)((ref.r + vg + vg_r) & 1023); px.b = (uint16_t)((ref.b + vg + vg_b) & 1023);
                        } else { rewind1(&R); rewind1(&R); px.r = px.g; px.b = px.g; }
                    } else if (op < 0xe0) {                                /* LUMA2 */
                        int vg = sx((uint32_t)((op & 31) << 2) | readBits10(&R, 2), 7);
                        px.g = (uint16_t)((ref.g + vg) & 1023);
                        if (!grey) {
                            int vg_r = sx(readBits10(&R, 6), 6), vg_b = sx(readBits10(&R, 6), 6);
                            px.r = (uint16_t)((ref.r + vg + vg_r) & 1023); px.b = (uint16_t)((ref.b + vg + vg_b) & 1023);
                        } else { px.r = px.g; px.b = px.g; }
                    } else if (op < 0xe8) {                                /* LUMA3 */
                        int vg = sx((uint32_t)((op & 7) << 6) | readBits10(&R, 6), 9);
                        px.g = (uint16_t)((ref.g + vg) & 1023);
                        if (!grey) {
                            int vg_r = sx(readBits10(&R, 8), 8), vg_b = sx(readBits10(&R, 8), 8);
                            px.r = (uint16_t)((ref.r + vg + vg_r) & 1023); px.b = (uint16_t)((ref.b + vg + vg_b) & 1023);
                        } else { px.r = px.g; px.b = px.g; }
                    } else if (op < 0xf0) {                                /* ADIFF */
                        int ad = sx((uint32_t)((op & 7) << 2) | readBits10(&R, 2), 5);
                        px.a = (uint16_t)((px.a + ad) & 1023); continue;
                    } else if ((op & 0xfc) == 0xf8) {                      /* ADIFF2 */
                        int ad = sx((uint32_t)((op & 3) << 6) | readBits10(&R, 6), 8);
                        px.a = (uint16_t)((px.a + ad) & 1023); continue;
                    } else if (op < 0xf8) {                                /* RUN */
                        run = op & 7;
                        if (run == 7) run = (int)readBits10(&R, 8) + 7;
                    } else if (op == 0xfd || op == 0xfe) {                 /* RGB / RGBA */
                        px.r = (uint16_t)readBits10(&R, 10);
                        if (!grey) { px.g = (uint16_t)readBits10(&R, 10); px.b = (uint16_t)readBits10(&R, 10); }
                        else { px.g = px.r; px.b = px.r; }
                        if (op == 0xfe) px.a = (uint16_t)readBits10(&R, 10);
                    } else if (op == 0xfc) { px.r = (uint16_t)readBits10(&R, 10); px.g = px.r; px.b = px.r; }   /* GRAY */
                    else finished = 1;                                     /* END */
                    break;
                }
                if (finished) break;
            }
            cur[x] = px;
        }
        if (finished) break;      /* goto finished: the row being decoded is never converted */
        uint16_t* line = (uint16_t*)(pixels + (size_t)desc->pitchBytes * y);
        for (int x = 0; x < W; ++x) {
            rgba10 q = cur[x];
            uint16_t r = (uint16_t)(q.r << 6 | (q.r >> 4)), g = (uint16_t)(q.g << 6 | (q.g >> 4));
            uint16_t b = (uint16_t)(q.b << 6 | (q.b >> 4)), a = (uint16_t)(q.a << 6 | (q.a >> 4));
            switch (channels) {
            default: case 4: line[x * 4] = r; line[x * 4 + 1] = g; line[x * 4 + 2] = b; line[x * 4 + 3] = a; break;
            case 3: line[x * 3] = r; line[x * 3 + 1] = g; line[x * 3 + 2] = b; break;
            case 2: line[x * 2] = r; line[x * 2 + 1] = a; break;
            case 1: line[x] = r; break;
            }
        }
        rgba10* t = cur; cur = last; last = t;
    }
    free(cur); free(last);
    return pixels;
}

/* ------------------------------------------------------------------------------------------ */
/* qoiplane_encode (qoiplane.d:109-375): 8-bit L / LA -> QOI-Plane stream (no LZ4 stage). Restated so that the
 * reference's round-trip property (image.d:2112-2183: encode -> decode == image) can be replayed for this codec and so
 * that the GPU encoder has a byte-exact target. */
typedef struct { uint8_t* bytes; int p; int hi; } nibw;
static void outputNibble(nibw* w, uint8_t nibble)                 /* :158-170 */
{
    if (w->hi) w->bytes[w->p] = (uint8_t)(nibble << 4); else w->bytes[w->p++] |= nibble;
    w->hi = !w->hi;
}
static void outputByte(nibw* w, uint8_t b)                        /* :172-184 */
{
    if (w->hi) w->bytes[w->p++] = b;
    else { w->bytes[w->p++] |= (uint8_t)(b >> 4); w->bytes[w->p] = (uint8_t)(b << 4); }
}
static void encodeRun(nibw* w, int* run)                          /* :186-212 */
{
    if (*run <= 3) outputNibble(w, (uint8_t)(0xc | (*run - 1)));
    else { *run -= 4; outputNibble(w, 0xf); outputByte(w, (uint8_t)*run); }
    *run = 0;
}
uint8_t* or_qoiplane_encode(const uint8_t* data, const or_qoix_desc* desc, int* out_len)
{
    if ((desc->channels != 1 && desc->channels != 2) || desc->width == 0 ||      /* width == 0 divides by zero in the reference */
        desc->height >= QOIX_PIXELS_MAX / desc->width || desc->compression != 0) return NULL;
    if (desc->bitdepth != 8) return NULL;
    const int channels = desc->channels;
    const int num_pixels = (int)(desc->width * desc->height);
    const int worst = channels == 1 ? 3 : 6;
    const int max_size = (num_pixels * worst + 1) / 2 + QOIX_HEADER_SIZE + 4;
    uint8_t* bytes = (uint8_t*)calloc((size_t)max_size + 2, 1);
    if (!bytes) return NULL;
    int p = 0;
    const uint32_t hdr[3] = {0x716F6978u, desc->width, desc->height};
    for (int k = 0; k < 3; ++k) { bytes[p++] = (uint8_t)(hdr[k] >> 24); bytes[p++] = (uint8_t)(hdr[k] >> 16); bytes[p++] = (uint8_t)(hdr[k] >> 8); bytes[p++] = (uint8_t)hdr[k]; }
    bytes[p++] = 1;
    bytes[p++] = desc->channels;
    bytes[p++] = desc->bitdepth;
    bytes[p++] = desc->colorspace;
    bytes[p++] = 0;
    uint32_t f[2];
    memcpy(&f[0], &desc->pixelAspectRatio, 4); memcpy(&f[1], &desc->resolutionY, 4);
    for (int k = 0; k < 2; ++k) { bytes[p++] = (uint8_t)(f[k] >> 24); bytes[p++] = (uint8_t)(f[k] >> 16); bytes[p++] = (uint8_t)(f[k] >> 8); bytes[p++] = (uint8_t)f[k]; }
    nibw W = {bytes, p, 1};
    uint8_t px_l = 0, px_a = 255, ref_l = 0, ref_a = 255;              /* initialPredictor (:95) */
    int run = 0, pixels_encoded = 0;
    for (int posy = 0; posy < (int)desc->height; ++posy) {
        const uint8_t* line = data + (ptrdiff_t)desc->pitchBytes * posy;
        const uint8_t* lineAbove = posy > 0 ? data + (ptrdiff_t)desc->pitchBytes * (posy - 1) : NULL;
        for (int posx = 0; posx < (int)desc->width; ++posx) {
            ref_l = px_l; ref_a = px_a;
            px_l = line[posx * channels];
            if (channels == 2) px_a = line[posx * channels + 1];
            if (px_l == ref_l && px_a == ref_a) {
                run++;
                if (run == 258 || pixels_encoded + 1 == num_pixels) encodeRun(&W, &run);
            } else {
                if (run > 0) encodeRun(&W, &run);
                const int8_t va = (int8_t)(px_a - ref_a);
                int colour = 1;
                if (va) {
                    if (va >= -7 && va <= 7) { outputNibble(&W, 0xb); outputNibble(&W, (uint8_t)(va + 8)); }
                    else { outputNibble(&W, 0xb); outputNibble(&W, 0x0); outputByte(&W, px_l); outputByte(&W, px_a); colour = 0; }
                }
                if (colour) {
                    const uint8_t px_top = posy > 0 ? lineAbove[posx * channels] : ref_l;
                    const uint8_t px_avg = (uint8_t)((px_top + ref_l + 1) / 2);
                    const int8_t diff_avg = (int8_t)(px_l - px_avg);
                    if (diff_avg >= -4 && diff_avg <= 3) outputNibble(&W, (uint8_t)(diff_avg + 4));
                    else if (diff_avg >= -16 && diff_avg <= 15) outputByte(&W, (uint8_t)(0x80 | (uint8_t)(diff_avg + 16)));
                    else { outputNibble(&W, 0xa); outputByte(&W, px_l); }
                }
            }
            pixels_encoded++;
        }
    }
    for (int i = 0; i < 9; ++i) outputNibble(&W, 0xf);                  /* :318-322 */
    if (!W.hi) outputNibble(&W, 0xf);
    *out_len = W.p;
    return bytes;
}

/* ------------------------------------------------------------------------------------------ */
/* qoix_encode (qoi2avg.d:376-617): rgb8 / rgba8 rows -> QOI2AVG stream (no LZ4 stage). Restated so that the reference's
 * round-trip property (image.d:2112-2183) can be replayed for this codec and so that the GPU encoder has a byte-exact
 * target. pitchBytes must be >= width * channels (the reference memcpy()s pitchBytes bytes of an rgba8 row, :459). */
static uint32_t q2_hash(uint32_t v) { return ((v * 2654435769u) >> 22) & 1023u; }                    /* :312-315 */
static uint32_t q2_v(rgba8 p) { return (uint32_t)p.r | (uint32_t)p.g << 8 | (uint32_t)p.b << 16 | (uint32_t)p.a << 24; }
uint8_t* or_qoix_encode(const uint8_t* data, const or_qoix_desc* desc, int* out_len)
{
    if (!data || !out_len || !desc || desc->width == 0 || desc->height == 0 || desc->channels < 3 || desc->channels > 4 ||
        desc->colorspace > 2 || desc->bitdepth != 8 || desc->compression != 0 || desc->height >= QOIX_PIXELS_MAX / desc->width) return NULL;
    const int W = (int)desc->width, H = (int)desc->height, channels = desc->channels;
    const int max_size = W * H * (channels + 1) + QOIX_HEADER_SIZE + 4;
    uint8_t* bytes = (uint8_t*)malloc((size_t)max_size + (size_t)W * 8);
    if (!bytes) return NULL;
    rgba8* inputScanline = (rgba8*)(bytes + max_size);
    rgba8* lastInputScanline = inputScanline + W;
    int p = 0;
    const uint32_t hdr[3] = {QOIX_MAGIC, desc->width, desc->height};
    for (int k = 0; k < 3; ++k) { bytes[p++] = (uint8_t)(hdr[k] >> 24); bytes[p++] = (uint8_t)(hdr[k] >> 16); bytes[p++] = (uint8_t)(hdr[k] >> 8); bytes[p++] = (uint8_t)hdr[k]; }
    bytes[p++] = 1; bytes[p++] = desc->channels; bytes[p++] = desc->bitdepth; bytes[p++] = desc->colorspace; bytes[p++] = 0;
    uint32_t f[2];
    memcpy(&f[0], &desc->pixelAspectRatio, 4); memcpy(&f[1], &desc->resolutionY, 4);
    for (int k = 0; k < 2; ++k) { bytes[p++] = (uint8_t)(f[k] >> 24); bytes[p++] = (uint8_t)(f[k] >> 16); bytes[p++] = (uint8_t)(f[k] >> 8); bytes[p++] = (uint8_t)f[k]; }
    uint8_t index_lookup[1024]; uint32_t index[64]; uint32_t index_pos = 0;
    memset(index, 0, sizeof(index)); memset(index_lookup, 0, sizeof(index_lookup));
    int run = 0, px_pos = 0;
    rgba8 px = {0, 0, 0, 255}, px_ref;
    const int px_end = W * H * channels - channels;
    for (int posy = 0; posy < H; ++posy) {
        const uint8_t* line = data + (ptrdiff_t)desc->pitchBytes * posy;
        for (int posx = 0; posx < W; ++posx) {                                                     /* :456-470 */
            if (channels == 4) { inputScanline[posx].r = line[posx * 4]; inputScanline[posx].g = line[posx * 4 + 1]; inputScanline[posx].b = line[posx * 4 + 2]; inputScanline[posx].a = line[posx * 4 + 3]; }
            else { inputScanline[posx].r = line[posx * 3]; inputScanline[posx].g = line[posx * 3 + 1]; inputScanline[posx].b = line[posx * 3 + 2]; inputScanline[posx].a = 255; }
        }
        for (int posx = 0; posx < W; ++posx) {
            px_ref = px;
            px = inputScanline[posx];
            if (q2_v(px) == q2_v(px_ref)) {
                run++;
                if (run == 1024 || px_pos == px_end) { run--; bytes[p++] = (uint8_t)(0xf8 | ((run >> 8) & 3)); bytes[p++] = (uint8_t)(run & 0xff); run = 0; }
            } else {
                const uint32_t hash = q2_hash(q2_v(px));
                if (run > 0) {
                    run--;
                    if (run < 8) bytes[p++] = (uint8_t)(0xf0 | run);
                    else { bytes[p++] = (uint8_t)(0xf8 | ((run >> 8) & 3)); bytes[p++] = (uint8_t)(run & 0xff); }
                    run = 0;
                }
                if (index[index_lookup[hash]] == q2_v(px)) bytes[p++] = (uint8_t)(0x80 | index_lookup[hash]);
                else {
                    index_lookup[hash] = (uint8_t)index_pos;
                    index[index_pos] = q2_v(px);
                    index_pos = (index_pos + 1) & 63;
                    const int8_t va = (int8_t)(px.a - px_ref.a);
                    int colour = 1;
                    if (va) {
                        if (va >= -4 && va <= 3) bytes[p++] = (uint8_t)(0xe8 | (va + 4));
                        else { bytes[p++] = 0xfe; bytes[p++] = px.r; bytes[p++] = px.g; bytes[p++] = px.b; bytes[p++] = px.a; colour = 0; }
                    }
                    if (colour) {
                        if (posy > 0) {
                            if (posx == 0) { px_ref.r = lastInputScanline[0].r; px_ref.g = lastInputScanline[0].g; px_ref.b = lastInputScanline[0].b; }
                            else {
                                const rgba8 a = px_ref, b = lastInputScanline[posx], c = lastInputScanline[posx - 1];
                                px_ref.r = (uint8_t)loco8(a.r, b.r, c.r); px_ref.g = (uint8_t)loco8(a.g, b.g, c.g); px_ref.b = (uint8_t)loco8(a.b, b.b, c.b);
                            }
                        }
                        const int8_t vg = (int8_t)(px.g - px_ref.g);
                        const int8_t vg_r = (int8_t)(px.r - px_ref.r - vg), vg_b = (int8_t)(px.b - px_ref.b - vg);
                        if (vg >= -4 && vg < 0 && vg_r >= -1 && vg_r <= 2 && vg_b >= -1 && vg_b <= 2)
                            bytes[p++] = (uint8_t)(0x00 | (vg + 4) << 4 | (vg_r + 1) << 2 | (vg_b + 1));
                        else if (vg >= 0 && vg <= 3 && vg_r >= -2 && vg_r <= 1 && vg_b >= -2 && vg_b <= 1)
                            bytes[p++] = (uint8_t)(0x00 | (vg + 4) << 4 | (vg_r + 2) << 2 | (vg_b + 2));
                        else if (px.g == px.r && px.g == px.b) { bytes[p++] = 0xfc; bytes[p++] = px.g; }
                        else if (vg_r >= -8 && vg_r <= 7 && vg >= -16 && vg <= 15 && vg_b >= -8 && vg_b <= 7) {
                            bytes[p++] = (uint8_t)(0xc0 | (vg + 16)); bytes[p++] = (uint8_t)((vg_r + 8) << 4 | (vg_b + 8));
                        } else if (vg_r >= -32 && vg_r <= 31 && vg >= -64 && vg <= 63 && vg_b >= -32 && vg_b <= 31) {
                            const int dv = ((vg + 64) << 12) | ((vg_r + 32) << 6) | (vg_b + 32);
                            bytes[p++] = (uint8_t)(0xe0 | ((dv >> 16) & 31)); bytes[p++] = (uint8_t)((dv >> 8) & 255); bytes[p++] = (uint8_t)(dv & 255);
                        } else { bytes[p++] = 0xfd; bytes[p++] = px.r; bytes[p++] = px.g; bytes[p++] = px.b; }
                    }
                }
            }
            px_pos += channels;
        }
        rgba8* t = inputScanline; inputScanline = lastInputScanline; lastInputScanline = t;
    }
    for (int i = 0; i < 4; ++i) bytes[p++] = 255;
    *out_len = p;
    return bytes;
}

/* ------------------------------------------------------------------------------------------ */
/* qoi10b_encode (qoi10b.d:136-500), stream version 1 (the only one the reference's encoder writes, :168): 16-bit
 * samples (10 significant bits) of 1-4 channels -> QOI-10b stream (no LZ4 stage). The encoder declares a colour index
 * (:233-235) but never emits an index opcode; the prediction is the rounded-up average of the previous pixel and the
 * pixel above (:357-362). */
typedef struct { uint8_t* bytes; int p; int currentBit; } bitw10;
static void outputBits10(bitw10* w, uint32_t x, int nbits)          /* :196-211 */
{
    for (int b = nbits - 2; b >= 0; b -= 2) {
        const uint8_t pair = (uint8_t)((x >> b) & 3);
        w->bytes[w->p] |= (uint8_t)(pair << (w->currentBit - 1));
        w->currentBit -= 2;
        if (w->currentBit == -1) { w->p++; w->bytes[w->p] = 0; w->currentBit = 7; }
    }
}
static int fits10(int v, int k) { return v >= 1024 - k || v < k; }
uint8_t* or_qoi10b_encode(const uint8_t* data, const or_qoix_desc* desc, int* out_len)
{
    if ((desc->channels != 1 && desc->channels != 2 && desc->channels != 3 && desc->channels != 4) || desc->width == 0 ||
        desc->height >= QOIX_PIXELS_MAX / desc->width || desc->compression != 0) return NULL;
    if (desc->bitdepth != 10) return NULL;
    const int channels = desc->channels, W = (int)desc->width, H = (int)desc->height;
    const int num_pixels = W * H;
    /* the reference sizes its buffer for 48 bits per pixel (:113, :152); ADIFF2 + RGB is 54, so allocate for 56 */
    const size_t max_size = ((size_t)num_pixels * 56 + 7) / 8 + QOIX_HEADER_SIZE + 5 + 2;
    uint8_t* bytes = (uint8_t*)calloc(max_size, 1);
    rgba10* cur = (rgba10*)calloc((size_t)W, sizeof(rgba10)), *last = (rgba10*)calloc((size_t)W, sizeof(rgba10));
    if (!bytes || !cur || !last) { free(bytes); free(cur); free(last); return NULL; }
    int p = 0;
    const uint32_t hdr[3] = {QOIX_MAGIC, desc->width, desc->height};
    for (int k = 0; k < 3; ++k) { bytes[p++] = (uint8_t)(hdr[k] >> 24); bytes[p++] = (uint8_t)(hdr[k] >> 16); bytes[p++] = (uint8_t)(hdr[k] >> 8); bytes[p++] = (uint8_t)hdr[k]; }
    bytes[p++] = 1; bytes[p++] = desc->channels; bytes[p++] = desc->bitdepth; bytes[p++] = desc->colorspace; bytes[p++] = 0;
    uint32_t f[2];
    memcpy(&f[0], &desc->pixelAspectRatio, 4); memcpy(&f[1], &desc->resolutionY, 4);
    for (int k = 0; k < 2; ++k) { bytes[p++] = (uint8_t)(f[k] >> 24); bytes[p++] = (uint8_t)(f[k] >> 16); bytes[p++] = (uint8_t)(f[k] >> 8); bytes[p++] = (uint8_t)f[k]; }
    bitw10 Wr = {bytes, p, 7};
    const int grey = channels == 1 || channels == 2;
    rgba10 px = {0, 0, 0, 1023}, ref;
    int run = 0, encoded = 0;
    for (int posy = 0; posy < H; ++posy) {
        const uint16_t* line = (const uint16_t*)(data + (ptrdiff_t)desc->pitchBytes * posy);
        for (int x = 0; x < W; ++x) {                                                          /* :254-297 */
            rgba10 q;
            if (channels == 4) { q.r = line[x * 4]; q.g = line[x * 4 + 1]; q.b = line[x * 4 + 2]; q.a = line[x * 4 + 3]; }
            else if (channels == 3) { q.r = line[x * 3]; q.g = line[x * 3 + 1]; q.b = line[x * 3 + 2]; q.a = 65535; }
            else if (channels == 2) { q.r = q.g = q.b = line[x * 2]; q.a = line[x * 2 + 1]; }
            else { q.r = q.g = q.b = line[x]; q.a = 65535; }
            q.r >>= 6; q.g >>= 6; q.b >>= 6; q.a >>= 6;
            cur[x] = q;
        }
        for (int x = 0; x < W; ++x) {
            ref = px;
            px = cur[x];
            if (px.r == ref.r && px.g == ref.g && px.b == ref.b && px.a == ref.a) {
                run++;
                if (run == 256 || encoded + 1 == num_pixels) goto flush_run;
                goto next;
            flush_run:
                run--;
                if (run < 7) outputBits10(&Wr, 0xf0u | (uint32_t)run, 8);
                else { outputBits10(&Wr, 0xf7u, 8); outputBits10(&Wr, (uint32_t)(run - 7), 8); }
                run = 0;
            } else {
                if (run > 0) {
                    run--;
                    if (run < 7) outputBits10(&Wr, 0xf0u | (uint32_t)run, 8);
                    else { outputBits10(&Wr, 0xf7u, 8); outputBits10(&Wr, (uint32_t)(run - 7), 8); }
                    run = 0;
                }
                const int va = (px.a - ref.a) & 1023;
                if (va) {
                    if (fits10(va, 16)) outputBits10(&Wr, (0x1du << 5) | (uint32_t)(va & 0x1f), 10);            /* QOI_OP_ADIFF */
                    else if (fits10(va, 128)) { outputBits10(&Wr, 0xf8u >> 2, 6); outputBits10(&Wr, (uint32_t)va, 8); }   /* QOI_OP_ADIFF2 */
                    else {
                        outputBits10(&Wr, 0xfeu, 8); outputBits10(&Wr, px.r, 10);                               /* QOI_OP_RGBA */
                        if (!grey) { outputBits10(&Wr, px.g, 10); outputBits10(&Wr, px.b, 10); }
                        outputBits10(&Wr, px.a, 10);
                        goto next;
                    }
                }
                if (posy > 0) {                                                                                /* version 1: average */
                    ref.r = (uint16_t)((ref.r + last[x].r + 1) >> 1);
                    ref.g = (uint16_t)((ref.g + last[x].g + 1) >> 1);
                    ref.b = (uint16_t)((ref.b + last[x].b + 1) >> 1);
                }
                const int vg = (px.g - ref.g) & 1023;
                const int vg_r = (px.r - ref.r - vg) & 1023, vg_b = (px.b - ref.b - vg) & 1023;
                if (fits10(vg_r, 4) && fits10(vg, 8) && fits10(vg_b, 4)) {
                    outputBits10(&Wr, 0x20u | (uint32_t)(vg & 0x0f), 6);                                        /* QOI_OP_LUMA0 */
                    if (!grey) outputBits10(&Wr, (uint32_t)((vg_r << 3) | (vg_b & 7)), 6);
                } else if (fits10(vg_r, 8) && fits10(vg, 16) && fits10(vg_b, 8)) {
                    outputBits10(&Wr, (uint32_t)(vg & 0x1f), 6);                                                /* QOI_OP_LUMA */
                    if (!grey) { outputBits10(&Wr, (uint32_t)vg_r, 4); outputBits10(&Wr, (uint32_t)vg_b, 4); }
                } else if (!grey && px.g == px.r && px.g == px.b) {
                    outputBits10(&Wr, 0xfcu, 8); outputBits10(&Wr, px.g, 10);                                   /* QOI_OP_GRAY */
                } else if (fits10(vg_r, 32) && fits10(vg, 64) && fits10(vg_b, 32)) {
                    outputBits10(&Wr, (0x6u << 7) | (uint32_t)(vg & 0x7f), 10);                                 /* QOI_OP_LUMA2 */
                    if (!grey) { outputBits10(&Wr, (uint32_t)vg_r, 6); outputBits10(&Wr, (uint32_t)vg_b, 6); }
                } else if (fits10(vg_r, 128) && fits10(vg, 256) && fits10(vg_b, 128)) {
                    outputBits10(&Wr, (0x1cu << 9) | (uint32_t)(vg & 0x1ff), 14);                               /* QOI_OP_LUMA3 */
                    if (!grey) { outputBits10(&Wr, (uint32_t)vg_r, 8); outputBits10(&Wr, (uint32_t)vg_b, 8); }
                } else {
                    outputBits10(&Wr, 0xfdu, 8); outputBits10(&Wr, px.r, 10);                                   /* QOI_OP_RGB */
                    if (!grey) { outputBits10(&Wr, px.g, 10); outputBits10(&Wr, px.b, 10); }
                }
            }
        next:
            encoded++;
        }
        rgba10* t = cur; cur = last; last = t;
    }
    for (int i = 0; i < 5; ++i) outputBits10(&Wr, 0xffu, 8);                                                    /* :488-491 */
    if (Wr.currentBit != 7) outputBits10(&Wr, 0xffu, Wr.currentBit + 1);
    free(cur); free(last);
    *out_len = Wr.p;
    return bytes;
}
