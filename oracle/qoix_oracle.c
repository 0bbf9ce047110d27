/*
 * qoix_oracle.c -- CPU restatement of QOI, the QOIX container + LZ4 wrapper and QOI-Plane10
 * (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * Follows: source/gamut/codecs/qoi.d:230-245,448-550 (qoi_decode), source/gamut/plugins/qoix.d:350-507
 * (qoix_lz4_decode, identifyTypeFromStream), source/gamut/codecs/qoi2avg.d:71-75,272-366 (header),
 * source/gamut/codecs/lz4.d:760-979 (LZ4_decompress_fast == LZ4_decompress_generic with
 * endOnOutputSize/withPrefix64k), source/gamut/codecs/qoiplane10.d:43-96,99-314 (encode),
 * :317-515 (decode).
 *
 * Restatement choices (equivalent on valid streams; the reference trusts its input):
 *  - every read past the end of the input is an error (LZ4) or yields 0xFF = END (opcode streams),
 *    where the reference reads out of bounds;
 *  - LZ4 match offsets that point before the start of the output are an error (the reference's
 *    "fast" variant does not check);
 *  - pixels after an END opcode are zero (the reference leaves malloc garbage).
 * The LZ4 *compressor* in here is a plain greedy hash matcher used only to synthesise test streams;
 * it is not a restatement of LZ4_compress (any valid LZ4 block decodes to the same bytes).
 *
 * parity: QOI is pinned by the QOI specification (PIL's independent QOI codec round-trips,
 * tests/test_oracle_qoix.py); QOIX/LZ4 are pinned by the reference's own round-trip property
 * (image.d:2112-2183 3x1 KAT, examples/qoix/source/main.d:113-121) through the restated encoder.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* QOI (qoi.d)                                                                                   */
#define QOI_MAGIC 0x716F6966u
#define QOI_HEADER_SIZE 14
#define QOI_PIXELS_MAX 400000000u
typedef struct { uint8_t r, g, b, a; } rgba_t;

static uint32_t rd32be(const uint8_t* b, int* p) { uint32_t a = b[(*p)++], c = b[(*p)++], d = b[(*p)++], e = b[(*p)++]; return a << 24 | c << 16 | d << 8 | e; }

/* qoi.d:448-550 */
uint8_t* or_qoi_decode(const uint8_t* bytes, int size, or_qoi_desc* desc, int channels)
{
    rgba_t index[64]; rgba_t px;
    int p = 0, run = 0;
    if ((channels != 0 && channels != 3 && channels != 4) || size < QOI_HEADER_SIZE + 8) return NULL;
    uint32_t magic = rd32be(bytes, &p);
    desc->width = rd32be(bytes, &p);
    desc->height = rd32be(bytes, &p);
    desc->channels = bytes[p++];
    desc->colorspace = bytes[p++];
    if (desc->width == 0 || desc->height == 0 || desc->channels < 3 || desc->channels > 4 || desc->colorspace > 1 ||
        magic != QOI_MAGIC || desc->height >= QOI_PIXELS_MAX / desc->width) return NULL;
    if (channels == 0) channels = desc->channels;
    int px_len = (int)(desc->width * desc->height * channels);
    uint8_t* pixels = (uint8_t*)malloc((size_t)px_len);
    if (!pixels) return NULL;
    memset(index, 0, sizeof(index));
    px.r = 0; px.g = 0; px.b = 0; px.a = 255;
    int chunks_len = size - 8;
    for (int px_pos = 0; px_pos < px_len; px_pos += channels) {
        if (run > 0) run--;
        else if (p < chunks_len) {
            int b1 = bytes[p++];
            if (b1 == 0xfe) { px.r = bytes[p++]; px.g = bytes[p++]; px.b = bytes[p++]; }
            else if (b1 == 0xff) { px.r = bytes[p++]; px.g = bytes[p++]; px.b = bytes[p++]; px.a = bytes[p++]; }
            else if ((b1 & 0xc0) == 0x00) px = index[b1];
            else if ((b1 & 0xc0) == 0x40) { px.r += ((b1 >> 4) & 3) - 2; px.g += ((b1 >> 2) & 3) - 2; px.b += (b1 & 3) - 2; }
            else if ((b1 & 0xc0) == 0x80) { int b2 = bytes[p++]; int vg = (b1 & 0x3f) - 32; px.r += vg - 8 + ((b2 >> 4) & 0x0f); px.g += vg; px.b += vg - 8 + (b2 & 0x0f); }
            else if ((b1 & 0xc0) == 0xc0) run = (b1 & 0x3f);
            index[(px.r * 3 + px.g * 5 + px.b * 7 + px.a * 11) % 64] = px;
        }
        pixels[px_pos + 0] = px.r; pixels[px_pos + 1] = px.g; pixels[px_pos + 2] = px.b;
        if (channels == 4) pixels[px_pos + 3] = px.a;
    }
    return pixels;
}


/* qoi.d:295-426 (qoi_encode): raw RGB / RGBA rows (pitch in bytes) -> QOI stream. */
uint8_t* or_qoi_encode(const uint8_t* data, uint32_t width, uint32_t height, int pitchBytes, int channels, int colorspace, int* out_len)
{
    rgba_t index[64]; rgba_t px, px_prev;
    if (!data || !out_len || width == 0 || height == 0 || channels < 3 || channels > 4 || colorspace < 0 || colorspace > 1 ||
        height >= QOI_PIXELS_MAX / width) return NULL;
    int max_size = (int)(width * height * (uint32_t)(channels + 1)) + QOI_HEADER_SIZE + 8;
    int p = 0;
    uint8_t* bytes = (uint8_t*)malloc((size_t)max_size);
    if (!bytes) return NULL;
    const uint32_t hdr[3] = {QOI_MAGIC, width, height};
    for (int k = 0; k < 3; ++k) { bytes[p++] = (uint8_t)(hdr[k] >> 24); bytes[p++] = (uint8_t)(hdr[k] >> 16); bytes[p++] = (uint8_t)(hdr[k] >> 8); bytes[p++] = (uint8_t)hdr[k]; }
    bytes[p++] = (uint8_t)channels;
    bytes[p++] = (uint8_t)colorspace;
    memset(index, 0, sizeof(index));
    int run = 0;
    px_prev.r = 0; px_prev.g = 0; px_prev.b = 0; px_prev.a = 255;
    px = px_prev;
    int px_len = (int)(width * height * (uint32_t)channels), px_end = px_len - channels, px_pos = 0;
    for (int posy = 0; posy < (int)height; ++posy) {
        const uint8_t* line = data + (size_t)pitchBytes * posy;
        for (int posx = 0; posx < (int)width; ++posx) {
            if (channels == 4) { px.r = line[posx * 4]; px.g = line[posx * 4 + 1]; px.b = line[posx * 4 + 2]; px.a = line[posx * 4 + 3]; }
            else { px.r = line[posx * 3]; px.g = line[posx * 3 + 1]; px.b = line[posx * 3 + 2]; }
            if (px.r == px_prev.r && px.g == px_prev.g && px.b == px_prev.b && px.a == px_prev.a) {
                run++;
                if (run == 62 || px_pos == px_end) { bytes[p++] = (uint8_t)(0xc0 | (run - 1)); run = 0; }
            } else {
                if (run > 0) { bytes[p++] = (uint8_t)(0xc0 | (run - 1)); run = 0; }
                int index_pos = (px.r * 3 + px.g * 5 + px.b * 7 + px.a * 11) % 64;
                if (index[index_pos].r == px.r && index[index_pos].g == px.g && index[index_pos].b == px.b && index[index_pos].a == px.a) {
                    bytes[p++] = (uint8_t)(0x00 | index_pos);
                } else {
                    index[index_pos] = px;
                    if (px.a == px_prev.a) {
                        int8_t vr = (int8_t)(px.r - px_prev.r), vg = (int8_t)(px.g - px_prev.g), vb = (int8_t)(px.b - px_prev.b);
                        int8_t vg_r = (int8_t)(vr - vg), vg_b = (int8_t)(vb - vg);
                        if (vr > -3 && vr < 2 && vg > -3 && vg < 2 && vb > -3 && vb < 2)
                            bytes[p++] = (uint8_t)(0x40 | (vr + 2) << 4 | (vg + 2) << 2 | (vb + 2));
                        else if (vg_r > -9 && vg_r < 8 && vg > -33 && vg < 32 && vg_b > -9 && vg_b < 8) {
                            bytes[p++] = (uint8_t)(0x80 | (vg + 32));
                            bytes[p++] = (uint8_t)((vg_r + 8) << 4 | (vg_b + 8));
                        } else { bytes[p++] = 0xfe; bytes[p++] = px.r; bytes[p++] = px.g; bytes[p++] = px.b; }
                    } else { bytes[p++] = 0xff; bytes[p++] = px.r; bytes[p++] = px.g; bytes[p++] = px.b; bytes[p++] = px.a; }
                }
            }
            px_prev = px;
            px_pos += channels;
        }
    }
    static const uint8_t padding[8] = {0, 0, 0, 0, 0, 0, 0, 1};
    for (int i = 0; i < 8; ++i) bytes[p++] = padding[i];
    *out_len = p;
    return bytes;
}

/* ------------------------------------------------------------------------------------------ */
/* LZ4 block (lz4.d:760-979)                                                                    */
enum { ML_BITS = 4, ML_MASK = 15, RUN_MASK = 15, MINMATCH = 4, COPYLENGTH = 8, LASTLITERALS = 5, MFLIMIT = 12 };

/* LZ4_decompress_fast (lz4.d:976): decode exactly originalSize output bytes. Returns the number of
 * input bytes read, or a negative value on error. `srcSize` bounds the reads (the reference has none). */
static int lz4_decompress_fast_bounded(const uint8_t* src, int srcSize, uint8_t* dst, int originalSize)
{
    const uint8_t* ip = src; const uint8_t* iend = src + srcSize;
    uint8_t* op = dst; uint8_t* oend = dst + originalSize;
    if (originalSize == 0) { if (srcSize < 1) return -1; return (*ip == 0) ? 1 : -1; }
    for (;;) {
        if (ip >= iend) return -1;
        unsigned token = *ip++;
        size_t length = token >> ML_BITS;
        if (length == RUN_MASK) {
            unsigned s;
            do { if (ip >= iend) return -1; s = *ip++; length += s; } while (s == 255);
        }
        uint8_t* cpy = op + length;
        if (length > (size_t)(oend - op)) return -1;
        if (cpy > oend - COPYLENGTH) {
            if (cpy != oend) return -1;                  /* block decoding must stop exactly there */
            if (length > (size_t)(iend - ip)) return -1;
            memcpy(op, ip, length);
            ip += length;
            break;
        }
        if (length > (size_t)(iend - ip)) return -1;
        memcpy(op, ip, length);
        ip += length; op = cpy;
        if (iend - ip < 2) return -1;
        size_t offset = (size_t)ip[0] | ((size_t)ip[1] << 8); ip += 2;
        if (offset == 0 || offset > (size_t)(op - dst)) return -1;   /* reference: unchecked */
        const uint8_t* match = op - offset;
        length = token & ML_MASK;
        if (length == ML_MASK) {
            unsigned s;
            do { if (ip >= iend) return -1; s = *ip++; length += s; } while (s == 255);
        }
        length += MINMATCH;
        if (length > (size_t)(oend - op)) return -1;
        cpy = op + length;
        if (cpy > oend - LASTLITERALS) return -1;        /* last LASTLITERALS bytes must be literals */
        while (op < cpy) *op++ = *match++;               /* byte-by-byte: overlap semantics */
    }
    return (int)(ip - src);
}
int or_lz4_decompress_fast(const uint8_t* src, uint8_t* dst, int originalSize)
{
    return lz4_decompress_fast_bounded(src, 0x7fffffff, dst, originalSize);
}
int or_lz4_compress_bound(int isize) { return isize + isize / 255 + 16; }      /* lz4.d:68 */

/* Test-stream generator: greedy LZ4 block compressor (valid per the block format's end conditions:
 * last 5 bytes literals, last match starts >= 12 bytes before the end). */
int or_lz4_compress(const uint8_t* src, uint8_t* dst, int n)
{
    enum { HB = 16 };
    int* table = (int*)malloc(sizeof(int) << HB);
    for (int i = 0; i < (1 << HB); ++i) table[i] = -1;
    uint8_t* op = dst;
    int anchor = 0, i = 0;
    const int mflimit = n - MFLIMIT;
    while (i < mflimit) {
        uint32_t v; memcpy(&v, src + i, 4);
        uint32_t h = (v * 2654435761u) >> (32 - HB);
        int cand = table[h];
        table[h] = i;
        uint32_t cv = 0;
        if (cand >= 0) memcpy(&cv, src + cand, 4);
        if (cand >= 0 && i - cand <= 65535 && cv == v) {
            int ml = 4;
            const int maxml = (n - LASTLITERALS) - i;
            while (ml < maxml && src[cand + ml] == src[i + ml]) ml++;
            int lit = i - anchor;
            uint8_t* token = op++;
            if (lit >= 15) { *token = 15 << 4; int l = lit - 15; while (l >= 255) { *op++ = 255; l -= 255; } *op++ = (uint8_t)l; }
            else *token = (uint8_t)(lit << 4);
            memcpy(op, src + anchor, (size_t)lit); op += lit;
            *op++ = (uint8_t)((i - cand) & 255); *op++ = (uint8_t)((i - cand) >> 8);
            int m = ml - 4;
            if (m >= 15) { *token |= 15; m -= 15; while (m >= 255) { *op++ = 255; m -= 255; } *op++ = (uint8_t)m; }
            else *token |= (uint8_t)m;
            i += ml; anchor = i;
        } else i++;
    }
    int lit = n - anchor;
    uint8_t* token = op++;
    if (lit >= 15) { *token = 15 << 4; int l = lit - 15; while (l >= 255) { *op++ = 255; l -= 255; } *op++ = (uint8_t)l; }
    else *token = (uint8_t)(lit << 4);
    memcpy(op, src + anchor, (size_t)lit); op += lit;
    free(table);
    return (int)(op - dst);
}

/* ------------------------------------------------------------------------------------------ */
/* QOIX header (qoi2avg.d:71-75,305-308)                                                         */
#define QOIX_MAGIC 0x716F6978u
#define QOIX_HEADER_SIZE 25
#define QOIX_PIXELS_MAX 400000000u
enum { OFF_VERSION = 12, OFF_CHANNELS = 13, OFF_BITDEPTH = 14, OFF_COLORSPACE = 15, OFF_COMPRESSION = 16 };

static void wr32be(uint8_t* b, int* p, uint32_t v) { b[(*p)++] = (uint8_t)(v >> 24); b[(*p)++] = (uint8_t)(v >> 16); b[(*p)++] = (uint8_t)(v >> 8); b[(*p)++] = (uint8_t)v; }
static float rd32f(const uint8_t* b, int* p) { uint32_t r = rd32be(b, p); float f; memcpy(&f, &r, 4); return f; }
static void wr32f(uint8_t* b, int* p, float f) { uint32_t r; memcpy(&r, &f, 4); wr32be(b, p, r); }

/* ------------------------------------------------------------------------------------------ */
/* QOI-Plane10 (qoiplane10.d)                                                                    */
typedef struct { uint16_t l, a; } la10_t;

/* qoiplane10.d:84-96 */
static int locoPredict(int left, int top, int topleft)
{
    int max_ab = left > top ? left : top;
    int min_ab = left < top ? left : top;
    if (topleft >= max_ab) return min_ab;
    else if (topleft <= min_ab) return max_ab;
    int d = left + top - topleft;
    if (d < 0) d = 0;
    if (d > 1023) d = 1023;
    return d;
}

typedef struct { uint8_t* bytes; int p; int currentBit; } bitw;
static void outputBits(bitw* w, uint32_t x, int nbits)      /* qoiplane10.d:139-155 */
{
    for (int b = nbits - 2; b >= 0; b -= 2) {
        uint8_t pair = (x >> b) & 3;
        w->bytes[w->p] |= (uint8_t)(pair << (w->currentBit - 1));
        w->currentBit -= 2;
        if (w->currentBit == -1) { w->p++; w->bytes[w->p] = 0; w->currentBit = 7; }
    }
}

/* qoiplane10.d:99-314 */
int or_test_loco_predict10(int left, int top, int topleft) { return locoPredict(left, top, topleft); }   /* test hook */

uint8_t* or_qoiplane10_encode(const uint8_t* data, const or_qoix_desc* desc, int* out_len)
{
    if ((desc->channels != 1 && desc->channels != 2) || desc->width == 0 || desc->height == 0 ||
        desc->height >= QOIX_PIXELS_MAX / desc->width || desc->compression != 0) return NULL;
    if (desc->bitdepth != 10) return NULL;
    int channels = desc->channels;
    int num_pixels = (int)(desc->width * desc->height);
    int worst_bits = (channels == 1) ? 14 : 28;
    int max_size = (int)(((long long)num_pixels * worst_bits + 7) / 8) + QOIX_HEADER_SIZE + 5 + 16;
    bitw w; w.p = 0; w.bytes = (uint8_t*)malloc((size_t)max_size);
    if (!w.bytes) return NULL;
    wr32be(w.bytes, &w.p, QOIX_MAGIC);
    wr32be(w.bytes, &w.p, desc->width);
    wr32be(w.bytes, &w.p, desc->height);
    w.bytes[w.p++] = 2;
    w.bytes[w.p++] = desc->channels;
    w.bytes[w.p++] = desc->bitdepth;
    w.bytes[w.p++] = desc->colorspace;
    w.bytes[w.p++] = 0;
    wr32f(w.bytes, &w.p, desc->pixelAspectRatio);
    wr32f(w.bytes, &w.p, desc->resolutionY);
    w.currentBit = 7; w.bytes[w.p] = 0;

    int run = 0, run1_pred = 0, run1_val = 0;
    la10_t px = {0, 1023}, px_ref = {0, 1023};
    int pixels_encoded = 0;
#define ENCODE_RUN() do { run--; if (run < 7) outputBits(&w, (0x6 << 3) | run, 6); else { outputBits(&w, (0x6 << 3) | 7, 6); outputBits(&w, run - 7, 8); } run = 0; } while (0)
#define FLUSH_RUN() do { int done_ = 0; if (run == 1) { int vg_ = (run1_val - run1_pred) & 1023; if (vg_ < 4 || vg_ >= (1024 - 4)) { outputBits(&w, vg_ & 0x07, 4); run = 0; done_ = 1; } } if (!done_) ENCODE_RUN(); } while (0)
    for (int posy = 0; posy < (int)desc->height; ++posy) {
        const uint16_t* line = (const uint16_t*)(data + (size_t)desc->pitchBytes * posy);
        const uint16_t* lineAbove = (posy > 0) ? (const uint16_t*)(data + (size_t)desc->pitchBytes * (posy - 1)) : NULL;
        for (int posx = 0; posx < (int)desc->width; ++posx) {
            px_ref = px;
            if (channels == 1) px.l = (uint16_t)(line[posx] >> 6);
            else { px.l = (uint16_t)(line[posx * 2] >> 6); px.a = (uint16_t)(line[posx * 2 + 1] >> 6); }
            int pred;
            if (posy == 0) pred = px_ref.l;
            else if (posx == 0) pred = lineAbove[0] >> 6;
            else pred = locoPredict(px_ref.l, lineAbove[posx * channels] >> 6, lineAbove[(posx - 1) * channels] >> 6);
            if (px.l == px_ref.l && px.a == px_ref.a) {
                if (run == 0) { run1_pred = pred; run1_val = px.l; }
                run++;
                if (run == 256 || (pixels_encoded + 1 == num_pixels)) FLUSH_RUN();
            } else {
                if (run > 0) FLUSH_RUN();
                int encoded = 0;
                int va = ((int)px.a - (int)px_ref.a) & 1023;
                if (va) {
                    if (va < 32 || va >= (1024 - 32)) outputBits(&w, (0x3e << 6) | (va & 0x3f), 12);
                    else { outputBits(&w, 0xfe, 8); outputBits(&w, px.l, 10); outputBits(&w, px.a, 10); encoded = 1; }
                }
                if (!encoded) {
                    int vg = ((int)px.l - pred) & 1023;
                    if (vg < 4 || vg >= (1024 - 4)) outputBits(&w, vg & 0x07, 4);
                    else if (vg < 32 || vg >= (1024 - 32)) outputBits(&w, 0x80 | (vg & 0x3f), 8);
                    else if (vg < 64 || vg >= (1024 - 64)) outputBits(&w, (0x1e << 7) | (vg & 0x7f), 12);
                    else outputBits(&w, (0xe << 10) | (vg & 0x3ff), 14);
                }
            }
            pixels_encoded++;
        }
    }
    for (int i = 0; i < 5; ++i) outputBits(&w, 255, 8);
    if (w.currentBit != 7) outputBits(&w, 0xff, w.currentBit + 1);
    *out_len = w.p;
    return w.bytes;
}

typedef struct { const uint8_t* bytes; int size; int p; int currentBit; } bitr;
/* reads past the end yield 1-bits (0xFF = END); the reference reads out of bounds */
static int read2Bits(bitr* r)                                 /* qoiplane10.d:377-387 */
{
    int byte = (r->p >= 0 && r->p < r->size) ? r->bytes[r->p] : 0xFF;
    int bit = (byte >> (r->currentBit - 1)) & 3;
    r->currentBit -= 2;
    if (r->currentBit == -1) { r->currentBit = 7; r->p++; }
    return bit;
}
static uint32_t readBits(bitr* r, int nbits) { uint32_t v = 0; for (int b = 0; b < nbits; b += 2) v = (v << 2) | (uint32_t)read2Bits(r); return v; }
static void rewindInputBit(bitr* r) { if (r->currentBit == 7) { r->p--; r->currentBit = -1; } r->currentBit++; }   /* :367-375 */

/* qoiplane10.d:317-515 */
uint8_t* or_qoiplane10_decode(const uint8_t* data, int size, or_qoix_desc* desc, int channels)
{
    if ((channels < 0 || channels > 2) || size < QOIX_HEADER_SIZE + 5) return NULL;
    const uint8_t* bytes = data;
    int p = 0;
    uint32_t magic = rd32be(bytes, &p);
    desc->width = rd32be(bytes, &p);
    desc->height = rd32be(bytes, &p);
    int qoix_version = bytes[p++];
    desc->channels = bytes[p++];
    desc->bitdepth = bytes[p++];
    desc->colorspace = bytes[p++];
    desc->compression = bytes[p++];
    desc->pixelAspectRatio = rd32f(bytes, &p);
    desc->resolutionY = rd32f(bytes, &p);
    if (desc->width == 0 || desc->height == 0 || desc->channels < 1 || desc->channels > 2 || desc->colorspace > 1 ||
        desc->bitdepth != 10 || qoix_version != 2 || desc->compression != 0 || magic != QOIX_MAGIC ||
        desc->height >= QOIX_PIXELS_MAX / desc->width) return NULL;
    if (channels == 0) channels = desc->channels;
    int stride = (int)desc->width * channels * 2;
    desc->pitchBytes = stride;
    int num_pixels = (int)(desc->width * desc->height);
    size_t output_bytes = (size_t)stride * desc->height;
    uint8_t* pixels = (uint8_t*)calloc(output_bytes ? output_bytes : 1, 1);
    if (!pixels) return NULL;
    bitr r = { bytes, size, p, 7 };
    la10_t px = {0, 1023}, px_ref = {0, 1023};
    int decoded_pixels = 0, run = 0;
    for (int posy = 0; posy < (int)desc->height; ++posy) {
        uint16_t* line = (uint16_t*)(pixels + (size_t)desc->pitchBytes * posy);
        const uint16_t* lineAbove = (posy > 0) ? (const uint16_t*)(pixels + (size_t)desc->pitchBytes * (posy - 1)) : NULL;
        for (int posx = 0; posx < (int)desc->width; ++posx) {
            px_ref = px;
            if (run > 0) run--;
            else if (decoded_pixels < num_pixels) {
                int pred;
                if (posy == 0) pred = px_ref.l;
                else if (posx == 0) pred = lineAbove[0] >> 6;
                else pred = locoPredict(px_ref.l, lineAbove[posx * channels] >> 6, lineAbove[(posx - 1) * channels] >> 6);
            decode_op: ;
                uint8_t op = (uint8_t)readBits(&r, 8);
                if (op < 0x80) {
                    int vg = (op >> 4) & 0x07; vg = (int)((uint32_t)vg << 29) >> 29;
                    rewindInputBit(&r); rewindInputBit(&r); rewindInputBit(&r); rewindInputBit(&r);
                    px.l = (uint16_t)((pred + vg) & 1023);
                } else if (op < 0xc0) {
                    int vg = op & 0x3f; vg = (int)((uint32_t)vg << 26) >> 26;
                    px.l = (uint16_t)((pred + vg) & 1023);
                } else if (op < 0xe0) {
                    run = (op >> 2) & 7;
                    rewindInputBit(&r); rewindInputBit(&r);
                    if (run == 7) run = (int)readBits(&r, 8) + 7;
                } else if (op < 0xf0) {
                    int vg = (int)(((op & 0x0f) << 6) | readBits(&r, 6)); vg = (int)((uint32_t)vg << 22) >> 22;
                    px.l = (uint16_t)((pred + vg) & 1023);
                } else if (op < 0xf8) {
                    int vg = (int)(((op & 0x07) << 4) | readBits(&r, 4)); vg = (int)((uint32_t)vg << 25) >> 25;
                    px.l = (uint16_t)((pred + vg) & 1023);
                } else if (op < 0xfc) {
                    int va = (int)(((op & 3) << 4) | readBits(&r, 4)); va = (int)((uint32_t)va << 26) >> 26;
                    px.a = (uint16_t)((px_ref.a + va) & 1023);
                    goto decode_op;
                } else if (op == 0xfe) {
                    px.l = (uint16_t)readBits(&r, 10);
                    px.a = (uint16_t)readBits(&r, 10);
                } else if (op == 0xff) {
                    goto finished;
                } else {
                    goto finished;      /* 0xfc / 0xfd reserved: assert(false) in the reference */
                }
                decoded_pixels++;
            }
            uint16_t l16 = (uint16_t)((px.l << 6) | (px.l >> 4));
            if (channels == 1) line[posx] = l16;
            else { uint16_t a16 = (uint16_t)((px.a << 6) | (px.a >> 4)); line[posx * 2] = l16; line[posx * 2 + 1] = a16; }
        }
    }
finished:
    return pixels;
}

/* ------------------------------------------------------------------------------------------ */
/* QOIX + LZ4 container (plugins/qoix.d)                                                         */
static int identifyTypeFromStream(int channels, int bitdepth, int premul, int* type)   /* qoix.d:476-507 */
{
    if (bitdepth == 8) {
        if (channels == 1) *type = OR_l8; else if (channels == 2) *type = premul ? OR_lap8 : OR_la8;
        else if (channels == 3) *type = OR_rgb8; else if (channels == 4) *type = premul ? OR_rgbap8 : OR_rgba8;
        else return 0;
    } else if (bitdepth == 10) {
        if (channels == 1) *type = OR_l16; else if (channels == 2) *type = premul ? OR_lap16 : OR_la16;
        else if (channels == 3) *type = OR_rgb16; else if (channels == 4) *type = premul ? OR_rgbap16 : OR_rgba16;
        else return 0;
    } else return 0;
    return 1;
}
static int validLoadFlags(int f)                                /* internals/types.d:563-578 */
{
    if ((f & 0x10000) && (f & 0x80000)) return 0;
    if ((f & 0x20000) && (f & 0x40000)) return 0;
    if ((f & 0x1000000) && (f & 0x2000000)) return 0;
    int n = 0; if (f & 0x100000) ++n; if (f & 0x200000) ++n; if (f & 0x400000) ++n;
    return n <= 1;
}
static int pixelTypeNumChannels(int t) { static const int c[18] = {1,1,1,2,2,2,2,2,2,3,3,3,4,4,4,4,4,4}; return c[t]; }

uint8_t* or_qoiplane_decode(const uint8_t* data, int size, or_qoix_desc* desc, int channels);
uint8_t* or_qoix_decode(const uint8_t* data, int size, or_qoix_desc* desc, int channels);
uint8_t* or_qoi10b_decode(const uint8_t* data, int size, or_qoix_desc* desc, int channels);

/* plugins/qoix.d:350-473 */
uint8_t* or_qoix_lz4_decode(const uint8_t* data, int size, or_qoix_desc* desc, int flags, int* decodedType)
{
    if (size < QOIX_HEADER_SIZE) return NULL;
    if (!validLoadFlags(flags)) return NULL;
    int compression = data[OFF_COMPRESSION], colorspace = data[OFF_COLORSPACE], streamChannels = data[OFF_CHANNELS];
    int streamBitdepth = data[OFF_BITDEPTH], streamVersion = data[OFF_VERSION];
    int streamType;
    if (!identifyTypeFromStream(streamChannels, streamBitdepth, colorspace == 2, &streamType)) return NULL;
    int uncompressedSize; const uint8_t* uncompressed = NULL; uint8_t* dec = NULL;
    if (compression == 1) {
        if (size < QOIX_HEADER_SIZE + 4) return NULL;
        int p = QOIX_HEADER_SIZE;
        int orig = (int)rd32be(data, &p);
        if (orig < 0) return NULL;
        dec = (uint8_t*)malloc((size_t)QOIX_HEADER_SIZE + (size_t)orig + 1);
        memcpy(dec, data, QOIX_HEADER_SIZE);
        dec[OFF_COMPRESSION] = 0;
        int qoilen = lz4_decompress_fast_bounded(data + QOIX_HEADER_SIZE + 4, size - QOIX_HEADER_SIZE - 4, dec + QOIX_HEADER_SIZE, orig);
        if (qoilen < 0) { free(dec); return NULL; }
        uncompressedSize = QOIX_HEADER_SIZE + orig;
        uncompressed = dec;
    } else if (compression == 0) { uncompressedSize = size; uncompressed = data; }
    else return NULL;
    uint8_t* image = NULL;
    *decodedType = streamType;
    int channels = pixelTypeNumChannels(streamType);
    if (streamBitdepth == 10) {
        if ((streamChannels == 1 || streamChannels == 2) && streamVersion >= 2) image = or_qoiplane10_decode(uncompressed, uncompressedSize, desc, channels);
        else image = or_qoi10b_decode(uncompressed, uncompressedSize, desc, channels);
    } else {
        if (streamChannels == 1 || streamChannels == 2) image = or_qoiplane_decode(uncompressed, uncompressedSize, desc, channels);
        else image = or_qoix_decode(uncompressed, uncompressedSize, desc, channels);
    }
    free(dec);
    return image;
}

/* plugins/qoix.d:251-339 with force_lz4: keep the LZ4 form even when it is not smaller (bench variant) */
uint8_t* or_qoix_lz4_encode(const uint8_t* pixels, const or_qoix_desc* desc, int force_lz4, int* out_len)
{
    int qoilen = 0; uint8_t* qoix = NULL;
    if (desc->bitdepth == 10 && (desc->channels == 1 || desc->channels == 2)) qoix = or_qoiplane10_encode(pixels, desc, &qoilen);
    else return NULL;   /* the other sub-encoders are not needed to synthesise the hot-path inputs */
    if (!qoix) return NULL;
    int datalen = qoilen - QOIX_HEADER_SIZE;
    int maxsize = or_lz4_compress_bound(datalen);
    uint8_t* lz4Data = (uint8_t*)malloc((size_t)QOIX_HEADER_SIZE + 4 + (size_t)maxsize);
    memcpy(lz4Data, qoix, QOIX_HEADER_SIZE);
    int p = QOIX_HEADER_SIZE;
    wr32be(lz4Data, &p, (uint32_t)datalen);
    int lz4Size = or_lz4_compress(qoix + QOIX_HEADER_SIZE, lz4Data + QOIX_HEADER_SIZE + 4, datalen);
    int useCompressed = force_lz4 || (lz4Size + 4 < datalen);
    if (useCompressed) {
        free(qoix);
        *out_len = QOIX_HEADER_SIZE + 4 + lz4Size;
        lz4Data[OFF_COMPRESSION] = 1;
        return lz4Data;
    }
    free(lz4Data);
    *out_len = qoilen;
    return qoix;
}

/* placeholders until the remaining sub-codecs are restated */
__attribute__((weak)) uint8_t* or_qoiplane_decode(const uint8_t* d, int s, or_qoix_desc* q, int c) { (void)d; (void)s; (void)q; (void)c; return NULL; }
__attribute__((weak)) uint8_t* or_qoix_decode(const uint8_t* d, int s, or_qoix_desc* q, int c) { (void)d; (void)s; (void)q; (void)c; return NULL; }
__attribute__((weak)) uint8_t* or_qoi10b_decode(const uint8_t* d, int s, or_qoix_desc* q, int c) { (void)d; (void)s; (void)q; (void)c; return NULL; }
