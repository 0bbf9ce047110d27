/*
 * tga_oracle.c -- CPU restatement of the reference's TGA decoder (TEST INFRASTRUCTURE ONLY; see oracle.h):
 *
 *   or_tga_load   TGADecoder.getImageInfo + decodeImage   source/gamut/codecs/tga.d:313-382, :384-588
 *                 stbi__tga_get_comp :590-617, stbi__tga_read_rgb16 :619-646
 *                 as loadTGA calls them (source/gamut/plugins/tga.d:45-105)
 *
 * The reference reads through an IOStream; Image.loadFromMemory gives it a MemoryFile (io.d:384-440), whose semantics
 * are restated by the cursor below: a read past the end fails, a seek may land ON the end but not after it.
 *
 * parity: pinned by PIL's independent TGA reader on the variants PIL can write (grey, RGB, RGBA, palette; raw and RLE;
 * both orientations), and by the formulas of the cited lines on hand-made files for what PIL cannot write (15/16-bit,
 * 16-bit palettes and indices, grey + alpha, ID field, palette start) -- tests/test_oracle_tga.py.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

typedef struct { const uint8_t* d; size_t len, p; } cur;
static int rd8(cur* c, int* err) { if (c->p + 1 > c->len) { *err = 1; return 0; } *err = 0; return c->d[c->p++]; }                 /* io.d:130 */
static int rd16(cur* c, int* err) { if (c->p + 2 > c->len) { *err = 1; return 0; } *err = 0; int v = c->d[c->p] | c->d[c->p + 1] << 8; c->p += 2; return v; } /* io.d:147 */
static int skip(cur* c, size_t n) { if (c->p + n > c->len) return 0; c->p += n; return 1; }                                          /* io.d:108, mseek :396 */
static int rdn(cur* c, uint8_t* dst, size_t n) { if (c->p + n > c->len) return 0; memcpy(dst, c->d + c->p, n); c->p += n; return 1; }

static int tga_get_comp(int bits_per_pixel, int is_grey, int* is_rgb16)      /* tga.d:590-617 */
{
    *is_rgb16 = 0;
    switch (bits_per_pixel) {
    case 8: return 1;
    case 16: if (is_grey) return 2; /* fallthrough */
    case 15: *is_rgb16 = 1; return 3;
    case 24: case 32: return bits_per_pixel / 8;
    default: return 0;
    }
}
static void tga_read_rgb16(cur* c, uint8_t* out, int* err)                   /* tga.d:619-646 */
{
    const int px = rd16(c, err);
    if (*err) return;
    const int r = (px >> 10) & 31, g = (px >> 5) & 31, b = px & 31;
    out[0] = (uint8_t)((r * 255) / 31); out[1] = (uint8_t)((g * 255) / 31); out[2] = (uint8_t)((b * 255) / 31);
}

/* Returns malloc()'d w*h*comp bytes (l8 / la8 / rgb8 / rgba8 by comp, plugins/tga.d:72-79) or NULL. */
uint8_t* or_tga_load(const uint8_t* data, size_t len, int* width, int* height, int* comp)
{
    cur C = {data, len, 0};
    int err = 0;
    /* ---- getImageInfo (tga.d:313-382) ---- */
    int paletteStart = 0, paletteLen = 0, cmapSize = 0;
    const int dataOffset = rd8(&C, &err); if (err) return NULL;
    const int cmapType = rd8(&C, &err); if (err || cmapType > 1) return NULL;
    int imageType = rd8(&C, &err); if (err) return NULL;
    if (cmapType == 1) {
        if (imageType != 1 && imageType != 9) return NULL;
        paletteStart = rd16(&C, &err); if (err) return NULL;
        paletteLen = rd16(&C, &err); if (err) return NULL;
        if (paletteLen == 0) return NULL;
        cmapSize = rd8(&C, &err); if (err) return NULL;
        if (cmapSize != 8 && cmapSize != 15 && cmapSize != 16 && cmapSize != 24 && cmapSize != 32) return NULL;
        if (!skip(&C, 4)) return NULL;
    } else {
        if (imageType != 2 && imageType != 3 && imageType != 10 && imageType != 11) return NULL;
        if (!skip(&C, 9)) return NULL;
    }
    const int W = rd16(&C, &err); if (err) return NULL;
    const int H = rd16(&C, &err); if (err) return NULL;
    if (W < 1 || H < 1) return NULL;
    const int bpp = rd8(&C, &err); if (err) return NULL;
    if (cmapType == 1 && bpp != 8 && bpp != 16) return NULL;
    if (bpp != 8 && bpp != 15 && bpp != 16 && bpp != 24 && bpp != 32) return NULL;
    /* ---- decodeImage (tga.d:384-588) ---- */
    int isRLE = 0;
    if (imageType >= 8) { imageType -= 8; isRLE = 1; }
    int inverted = rd8(&C, &err); if (err) return NULL;
    inverted = 1 - ((inverted >> 5) & 1);
    const int isIndexed = cmapType != 0;
    int rgb16 = 0;
    const int components = isIndexed ? tga_get_comp(cmapSize, 0, &rgb16) : tga_get_comp(bpp, imageType == 3, &rgb16);
    if (!skip(&C, (size_t)dataOffset)) return NULL;
    const long long allocationSize = (long long)W * H * components;       /* <= 65535^2 * 4: far below GAMUT_MAX_IMAGE_BYTES */
    uint8_t* out = (uint8_t*)malloc((size_t)allocationSize);
    if (!out) return NULL;
    uint8_t* palette = NULL;
    if (!isIndexed && !isRLE && !rgb16) {
        for (int i = 0; i < H; ++i) {
            const int row = inverted ? H - i - 1 : i;
            if (!rdn(&C, out + (size_t)row * W * components, (size_t)W * components)) { free(out); return NULL; }
        }
    } else {
        if (isIndexed) {
            if (!skip(&C, (size_t)paletteStart)) { free(out); return NULL; }
            palette = (uint8_t*)malloc((size_t)paletteLen * components);
            if (rgb16) {
                for (int i = 0; i < paletteLen; ++i) { tga_read_rgb16(&C, palette + (size_t)i * components, &err); if (err) { free(palette); free(out); return NULL; } }
            } else if (!rdn(&C, palette, (size_t)paletteLen * components)) { free(palette); free(out); return NULL; }
        }
        int RLE_count = 0, RLE_repeating = 0, read_next_pixel = 1;
        uint8_t raw[4] = {0, 0, 0, 0};
        for (int i = 0; i < W * H; ++i) {
            if (isRLE) {
                if (RLE_count == 0) {
                    const int cmd = rd8(&C, &err); if (err) goto bad;
                    RLE_count = 1 + (cmd & 127); RLE_repeating = cmd >> 7; read_next_pixel = 1;
                } else if (!RLE_repeating) read_next_pixel = 1;
            } else read_next_pixel = 1;
            if (read_next_pixel) {
                if (isIndexed) {
                    int idx = bpp == 8 ? rd8(&C, &err) : rd16(&C, &err);
                    if (err) goto bad;
                    if (idx >= paletteLen) idx = 0;
                    for (int j = 0; j < components; ++j) raw[j] = palette[idx * components + j];
                } else if (rgb16) { tga_read_rgb16(&C, raw, &err); if (err) goto bad; }
                else for (int j = 0; j < components; ++j) { raw[j] = (uint8_t)rd8(&C, &err); if (err) goto bad; }
                read_next_pixel = 0;
            }
            for (int j = 0; j < components; ++j) out[(size_t)i * components + j] = raw[j];
            --RLE_count;
        }
        if (inverted)
            for (int j = 0; j * 2 < H; ++j) {
                uint8_t* a = out + (size_t)j * W * components; uint8_t* b = out + (size_t)(H - 1 - j) * W * components;
                for (int i = 0; i < W * components; ++i) { const uint8_t t = a[i]; a[i] = b[i]; b[i] = t; }
            }
        free(palette);
    }
    if (components >= 3 && !rgb16)
        for (int i = 0; i < W * H; ++i) { uint8_t* p = out + (size_t)i * components; const uint8_t t = p[0]; p[0] = p[2]; p[2] = t; }
    *width = W; *height = H; *comp = components;
    return out;
bad:
    free(palette); free(out);
    return NULL;
}

/* ------------------------------------------------------------------------------------------ */
/* saveTGA (plugins/tga.d:123-149) -> TGAEncoder.initialize / encodeScanline (codecs/tga.d:62-292) with RLE enabled:
 * l8 / la8 / rgb8 / rgba8 rows (type = PixelType value 0, 3, 9, 12) -> a 24- or 32-bit run-length TGA, bottom-up.
 * `data` = first scanline, pitchBytes signed. Returns malloc()'d file bytes or NULL (unsupported type, side > 65535). */
uint8_t* or_tga_encode(const uint8_t* data, int type, int width, int height, int pitchBytes, int* out_len)
{
    int channels;
    if (width > 65535 || height > 65535 || width < 0 || height < 0) return NULL;
    switch (type) {                                             /* :88-111 */
    case 0: channels = 3; break;      /* l8    -> rgb8 */
    case 3: channels = 4; break;      /* la8   -> rgba8 */
    case 9: channels = 3; break;      /* rgb8 */
    case 12: channels = 4; break;     /* rgba8 */
    default: return NULL;
    }
    const size_t cap = 18 + (size_t)height * ((size_t)width * (channels + 1) + 2);
    uint8_t* out = (uint8_t*)malloc(cap);
    uint8_t* scanSpace = (uint8_t*)malloc((size_t)width * channels + (size_t)width * 2 + 1);
    if (!out || !scanSpace) { free(out); free(scanSpace); return NULL; }
    int8_t* similarMask = (int8_t*)(scanSpace + (size_t)width * channels);
    int8_t* opcode = similarMask + width;
    size_t p = 0;
    memset(out, 0, 18);                                         /* header :120-131 */
    out[2] = 10;
    out[12] = (uint8_t)(width & 0xff); out[13] = (uint8_t)((width & 0xff00) >> 8);
    out[14] = (uint8_t)(height & 0xff); out[15] = (uint8_t)((height & 0xff00) >> 8);
    out[16] = (uint8_t)(channels * 8);
    p = 18;
    for (int y = height - 1; y >= 0; --y) {                     /* plugins/tga.d:141-145 */
        const uint8_t* scan = data + (ptrdiff_t)pitchBytes * y;
        if (width == 0) continue;
        /* _scanConvert, then swap R and B (:152-178) */
        for (int x = 0; x < width; ++x) {
            uint8_t r, g, b, a = 255;
            if (type == 0) r = g = b = scan[x];
            else if (type == 3) { r = g = b = scan[2 * x]; a = scan[2 * x + 1]; }
            else if (type == 9) { r = scan[3 * x]; g = scan[3 * x + 1]; b = scan[3 * x + 2]; }
            else { r = scan[4 * x]; g = scan[4 * x + 1]; b = scan[4 * x + 2]; a = scan[4 * x + 3]; }
            scanSpace[channels * x] = b; scanSpace[channels * x + 1] = g; scanSpace[channels * x + 2] = r;
            if (channels == 4) scanSpace[4 * x + 3] = a;
        }
        {   /* 1. similarity between consecutive pixels (:186-203) */
            uint8_t last[4] = {0, 0, 0, 0};
            for (int x = 0; x < width; ++x) {
                uint8_t c[4] = {scanSpace[channels * x], scanSpace[channels * x + 1], scanSpace[channels * x + 2],
                                (uint8_t)(channels == 3 ? 255 : scanSpace[x * 4 + 3])};
                similarMask[x] = memcmp(c, last, 4) == 0 ? 1 : 0;
                memcpy(last, c, 4);
            }
            similarMask[0] = 0;
        }
        int numSame = 0, numDifferent = 0;                       /* 2. (:207-243) */
        for (int x = width - 1; x >= 0; --x) {
            const float bppRaw = (1 + numDifferent * channels) / (float)numDifferent;
            const float bppRLE = (1 + channels) / (float)numSame;
            if (bppRaw <= bppRLE) opcode[x] = (int8_t)numDifferent;
            else opcode[x] = (int8_t)(0x80 | numSame);
            if (similarMask[x]) { numSame += 1; if (numSame >= 127) numSame = 127; numDifferent = 0; }
            else { numDifferent += 1; if (numDifferent >= 127) numDifferent = 127; numSame = 0; }
        }
        for (int x = 0; x < width; ) {                           /* 3. (:252-271) */
            const int8_t hint = opcode[x];
            out[p++] = (uint8_t)hint;
            const int num = (hint & 127) + 1;
            const size_t nbytes = hint >= 0 ? (size_t)num * channels : (size_t)channels;
            memcpy(out + p, &scanSpace[x * channels], nbytes);
            p += nbytes;
            x += num;
        }
    }
    free(scanSpace);
    *out_len = (int)p;
    return out;
}
