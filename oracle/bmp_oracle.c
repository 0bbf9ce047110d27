/*
 * bmp_oracle.c -- CPU restatement of the reference's BMP path and of its format detection (TEST INFRASTRUCTURE ONLY:
 * used by tests/, smoke() and bench.py's cpu legs; the product never links it).
 *
 * Follows source/gamut/codecs/stbdec.d (the stb_image 2.29 BMP loader, :2112-2510) function by function, including
 * the stbi__context buffering of a callback stream (:461-503, :780-842) -- the 128-byte refill buffer decides what a
 * negative stbi__skip does -- and the detect procs of plugins/ (one per format) + Image.identifyFormatFromStream (image.d:1045-1061).
 * Pinned by the reference's KAT for issue67.bmp (examples/test-suite/source/main.d:161-170: 32x32, 200 x 100 dpi, pixel
 * aspect ratio 2) and by PIL's independent BMP codec on every bit depth it writes (tests/test_oracle_bmp.py).
 * Divergence (corrupt input only): palette entries beyond the declared palette size read as zero (the reference reads
 * an uninitialised stack array, :2268).
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

/* ---- stbi__context over a memory "callback" stream (stbdec.d:461-503, 780-842) ---- */
typedef struct {
    const uint8_t* data; size_t len;
    size_t stream_pos;                  /* what the io callbacks have consumed */
    uint8_t buf[128]; int buf_n, buf_cur;
    int read_from_callbacks, callback_already_read;
    uint32_t img_x, img_y; int img_n;
    float ppmX, ppmY, pixelAspectRatio;
} sctx;

static void refill(sctx* s)                                  /* stbi__refill_buffer :780-795 */
{
    size_t n = s->len - s->stream_pos; if (n > 128) n = 128;
    memcpy(s->buf, s->data + s->stream_pos, n);
    s->stream_pos += n;
    s->callback_already_read += s->buf_cur;                  /* img_buffer - img_buffer_original */
    if (n == 0) { s->read_from_callbacks = 0; s->buf_cur = 0; s->buf_n = 1; s->buf[0] = 0; }
    else { s->buf_cur = 0; s->buf_n = (int)n; }
}
static void start(sctx* s, const uint8_t* d, size_t len)     /* stbi__start_callbacks :484-494 */
{
    memset(s, 0, sizeof(*s));
    s->data = d; s->len = len; s->read_from_callbacks = 1;
    refill(s);
}
static int get8(sctx* s)                                      /* stbi__get8 :797-806 */
{
    if (s->buf_cur < s->buf_n) return s->buf[s->buf_cur++];
    if (s->read_from_callbacks) { refill(s); return s->buf[s->buf_cur++]; }
    return 0;
}
static int get16le(sctx* s) { int z = get8(s); return z + (get8(s) << 8); }                      /* :882 */
static uint32_t get32le(sctx* s) { uint32_t z = (uint32_t)get16le(s); z += (uint32_t)get16le(s) << 16; return z; }  /* :888 */
static void skip(sctx* s, int n)                              /* stbi__skip :822-842 */
{
    if (n == 0) return;
    if (n < 0) { s->buf_cur = s->buf_n; return; }
    int blen = s->buf_n - s->buf_cur;
    if (blen < n) {
        s->buf_cur = s->buf_n;
        size_t adv = (size_t)(n - blen);                      /* io.skip: a seek of the memory stream */
        s->stream_pos = s->stream_pos + adv > s->len ? s->len : s->stream_pos + adv;
        return;
    }
    s->buf_cur += n;
}

typedef struct { int bpp, offset, hsz; uint32_t mr, mg, mb, ma, all_a; int extra_read; } bmp_data;

static int set_mask_defaults(bmp_data* info, int compress)   /* stbi__bmp_set_mask_defaults :2121-2145 */
{
    if (compress == 3) return 1;
    if (compress == 0) {
        if (info->bpp == 16) { info->mr = 31u << 10; info->mg = 31u << 5; info->mb = 31u << 0; }
        else if (info->bpp == 32) { info->mr = 0xffu << 16; info->mg = 0xffu << 8; info->mb = 0xffu << 0; info->ma = 0xffu << 24; info->all_a = 0; }
        else info->mr = info->mg = info->mb = info->ma = 0;
        return 1;
    }
    return 0;
}

static int parse_header(sctx* s, bmp_data* info)              /* stbi__bmp_parse_header :2147-2239 */
{
    int hsz;
    if (get8(s) != 'B' || get8(s) != 'M') return 0;
    get32le(s); get16le(s); get16le(s);
    info->offset = (int)get32le(s);
    info->hsz = hsz = (int)get32le(s);
    info->mr = info->mg = info->mb = info->ma = 0;
    info->extra_read = 14;
    s->ppmX = -1; s->ppmY = -1; s->pixelAspectRatio = -1;
    if (info->offset < 0) return 0;
    if (hsz != 12 && hsz != 40 && hsz != 56 && hsz != 108 && hsz != 124) return 0;
    if (hsz == 12) { s->img_x = (uint32_t)get16le(s); s->img_y = (uint32_t)get16le(s); }
    else { s->img_x = get32le(s); s->img_y = get32le(s); }
    if (get16le(s) != 1) return 0;
    info->bpp = get16le(s);
    if (hsz != 12) {
        int compress = (int)get32le(s);
        if (compress == 1 || compress == 2) return 0;
        if (compress >= 4) return 0;
        if (compress == 3 && info->bpp != 16 && info->bpp != 32) return 0;
        get32le(s);
        int xppm = (int)get32le(s), yppm = (int)get32le(s);
        if (xppm > 1) s->ppmX = (float)xppm;
        if (yppm > 1) s->ppmY = (float)yppm;
        if (s->ppmX != -1 && s->ppmY != -1) s->pixelAspectRatio = s->ppmX / s->ppmY;
        get32le(s); get32le(s);
        if (hsz == 40 || hsz == 56) {
            if (hsz == 56) { get32le(s); get32le(s); get32le(s); get32le(s); }
            if (info->bpp == 16 || info->bpp == 32) {
                if (compress == 0) set_mask_defaults(info, compress);
                else if (compress == 3) {
                    info->mr = get32le(s); info->mg = get32le(s); info->mb = get32le(s);
                    info->extra_read += 12;
                    if (info->mr == info->mg && info->mg == info->mb) return 0;
                } else return 0;
            }
        } else {
            if (hsz != 108 && hsz != 124) return 0;
            info->mr = get32le(s); info->mg = get32le(s); info->mb = get32le(s); info->ma = get32le(s);
            if (compress != 3) set_mask_defaults(info, compress);
            get32le(s);
            for (int i = 0; i < 12; ++i) get32le(s);
            if (hsz == 124) { get32le(s); get32le(s); get32le(s); get32le(s); }
        }
    }
    return 1;
}

static int high_bit(uint32_t z)                               /* stbi__high_bit :2468-2478 */
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
static int bitcount(uint32_t a)                               /* stbi__bitcount :2480-2488 */
{
    a = (a & 0x55555555) + ((a >> 1) & 0x55555555);
    a = (a & 0x33333333) + ((a >> 2) & 0x33333333);
    a = (a + (a >> 4)) & 0x0f0f0f0f;
    a = (a + (a >> 8));
    a = (a + (a >> 16));
    return (int)(a & 0xff);
}
static int shiftsigned(uint32_t v, int shift, int bits)       /* stbi__shiftsigned :2493-2512 */
{
    static const uint32_t mul_table[9] = {0, 0xff, 0x55, 0x49, 0x11, 0x21, 0x41, 0x81, 0x01};
    static const uint32_t shift_table[9] = {0, 0, 0, 1, 0, 2, 4, 6, 0};
    if (shift < 0) v <<= -shift; else v >>= shift;
    v >>= (8 - bits);
    return (int)((uint32_t)v * mul_table[bits]) >> shift_table[bits];
}
static uint8_t compute_y(int r, int g, int b) { return (uint8_t)(((r * 77) + (g * 150) + (29 * b)) >> 8); }   /* :911 */

/* stbi__mad3sizes_valid(a, b, c, 0) (:553-590): the products must fit an int */
static int mad3_valid(int a, int b, int c)
{
    if (a < 0 || b < 0 || c < 0) return 0;
    if (b && a > 0x7fffffff / b) return 0;
    long long ab = (long long)a * b;
    if (c && ab > 0x7fffffff / c) return 0;
    return 1;
}

/* stbi_load_from_callbacks -> stbi__load_main -> stbi__bmp_load (:613-633, :2263-2466) on a memory stream */
uint8_t* or_bmp_load(const uint8_t* data, size_t len, int req_comp, int* x, int* y, int* comp,
                     float* ppmX, float* ppmY, float* pixelRatio)
{
    sctx S; sctx* s = &S;
    start(s, data, len);
    /* stbi__bmp_test (:2241-2261) + stbi__rewind: the test looks at the first buffer only */
    {
        sctx t = S;
        int ok = get8(&t) == 'B' && get8(&t) == 'M';
        if (ok) { get32le(&t); get16le(&t); get16le(&t); get32le(&t); int sz = (int)get32le(&t); ok = sz == 12 || sz == 40 || sz == 56 || sz == 108 || sz == 124; }
        if (!ok) return NULL;
    }
    uint32_t mr, mg, mb, ma, all_a;
    uint8_t pal[256][4];
    memset(pal, 0, sizeof(pal));
    int psize = 0, i, j, width, pad, target;
    bmp_data info; memset(&info, 0, sizeof(info));
    info.all_a = 255;
    if (!parse_header(s, &info)) return NULL;
    const int flip_vertically = ((int)s->img_y) > 0;
    { int iy = (int)s->img_y; s->img_y = (uint32_t)(iy < 0 ? -iy : iy); }
    if (s->img_y > (1u << 24) || s->img_x > (1u << 24)) return NULL;
    mr = info.mr; mg = info.mg; mb = info.mb; ma = info.ma; all_a = info.all_a;
    if (info.hsz == 12) { if (info.bpp < 24) psize = (info.offset - info.extra_read - 24) / 3; }
    else { if (info.bpp < 16) psize = (info.offset - info.extra_read - info.hsz) >> 2; }
    if (psize == 0) {
        int bytes_read_so_far = s->callback_already_read + s->buf_cur;
        if (bytes_read_so_far <= 0 || bytes_read_so_far > 1024) return NULL;
        if (info.offset < bytes_read_so_far || info.offset - bytes_read_so_far > 256 * 4) return NULL;
        skip(s, info.offset - bytes_read_so_far);
    }
    if (info.bpp == 24 && ma == 0xff000000u) s->img_n = 3; else s->img_n = ma ? 4 : 3;
    target = (req_comp && req_comp >= 3) ? req_comp : s->img_n;
    if (!mad3_valid(target, (int)s->img_x, (int)s->img_y)) return NULL;
    uint8_t* out = (uint8_t*)malloc((size_t)target * s->img_x * s->img_y + 1);
    if (!out) return NULL;
    const int W = (int)s->img_x, H = (int)s->img_y;
    if (info.bpp < 16) {
        size_t z = 0;
        if (psize == 0 || psize > 256) { free(out); return NULL; }
        for (i = 0; i < psize; ++i) {
            pal[i][2] = (uint8_t)get8(s); pal[i][1] = (uint8_t)get8(s); pal[i][0] = (uint8_t)get8(s);
            if (info.hsz != 12) get8(s);
            pal[i][3] = 255;
        }
        skip(s, info.offset - info.extra_read - info.hsz - psize * (info.hsz == 12 ? 3 : 4));
        if (info.bpp == 1) width = (W + 7) >> 3;
        else if (info.bpp == 4) width = (W + 1) >> 1;
        else if (info.bpp == 8) width = W;
        else { free(out); return NULL; }
        pad = (-width) & 3;
        if (info.bpp == 1) {
            for (j = 0; j < H; ++j) {
                int bit_offset = 7, v = get8(s);
                for (i = 0; i < W; ++i) {
                    int color = (v >> bit_offset) & 0x1;
                    out[z++] = pal[color][0]; out[z++] = pal[color][1]; out[z++] = pal[color][2];
                    if (target == 4) out[z++] = 255;
                    if (i + 1 == W) break;
                    if ((--bit_offset) < 0) { bit_offset = 7; v = get8(s); }
                }
                skip(s, pad);
            }
        } else {
            for (j = 0; j < H; ++j) {
                for (i = 0; i < W; i += 2) {
                    int v = get8(s), v2 = 0;
                    if (info.bpp == 4) { v2 = v & 15; v >>= 4; }
                    out[z++] = pal[v][0]; out[z++] = pal[v][1]; out[z++] = pal[v][2];
                    if (target == 4) out[z++] = 255;
                    if (i + 1 == W) break;
                    v = (info.bpp == 8) ? get8(s) : v2;
                    out[z++] = pal[v][0]; out[z++] = pal[v][1]; out[z++] = pal[v][2];
                    if (target == 4) out[z++] = 255;
                }
                skip(s, pad);
            }
        }
    } else {
        int rshift = 0, gshift = 0, bshift = 0, ashift = 0, rcount = 0, gcount = 0, bcount = 0, acount = 0;
        size_t z = 0;
        int easy = 0;
        skip(s, info.offset - info.extra_read - info.hsz);
        if (info.bpp == 24) width = 3 * W;
        else if (info.bpp == 16) width = 2 * W;
        else width = 0;
        pad = (-width) & 3;
        if (info.bpp == 24) easy = 1;
        else if (info.bpp == 32) { if (mb == 0xff && mg == 0xff00 && mr == 0x00ff0000 && ma == 0xff000000u) easy = 2; }
        if (!easy) {
            if (!mr || !mg || !mb) { free(out); return NULL; }
            rshift = high_bit(mr) - 7; rcount = bitcount(mr);
            gshift = high_bit(mg) - 7; gcount = bitcount(mg);
            bshift = high_bit(mb) - 7; bcount = bitcount(mb);
            ashift = high_bit(ma) - 7; acount = bitcount(ma);
            if (rcount > 8 || gcount > 8 || bcount > 8 || acount > 8) { free(out); return NULL; }
        }
        for (j = 0; j < H; ++j) {
            if (easy) {
                for (i = 0; i < W; ++i) {
                    uint8_t a;
                    out[z + 2] = (uint8_t)get8(s); out[z + 1] = (uint8_t)get8(s); out[z + 0] = (uint8_t)get8(s);
                    z += 3;
                    a = (uint8_t)(easy == 2 ? get8(s) : 255);
                    all_a |= a;
                    if (target == 4) out[z++] = a;
                }
            } else {
                const int bpp = info.bpp;
                for (i = 0; i < W; ++i) {
                    uint32_t v = (bpp == 16 ? (uint32_t)get16le(s) : get32le(s));
                    uint32_t a;
                    out[z++] = (uint8_t)(shiftsigned(v & mr, rshift, rcount) & 255);
                    out[z++] = (uint8_t)(shiftsigned(v & mg, gshift, gcount) & 255);
                    out[z++] = (uint8_t)(shiftsigned(v & mb, bshift, bcount) & 255);
                    a = (ma ? (uint32_t)shiftsigned(v & ma, ashift, acount) : 255u);
                    all_a |= a;
                    if (target == 4) out[z++] = (uint8_t)(a & 255);
                }
            }
            skip(s, pad);
        }
    }
    if (target == 4 && all_a == 0) for (long long k = 4LL * W * H - 1; k >= 0; k -= 4) out[k] = 255;
    if (flip_vertically) {
        for (j = 0; j < H >> 1; ++j) {
            uint8_t* p1 = out + (size_t)j * W * target;
            uint8_t* p2 = out + (size_t)(H - 1 - j) * W * target;
            for (i = 0; i < W * target; ++i) { uint8_t t = p1[i]; p1[i] = p2[i]; p2[i] = t; }
        }
    }
    if (req_comp && req_comp != target) {
        /* stbi__convert_format (:916-1054) for the cases that can occur here: 3|4 -> 1|2 */
        uint8_t* good = (uint8_t*)malloc((size_t)req_comp * W * H + 1);
        for (size_t k = 0; k < (size_t)W * H; ++k) {
            const uint8_t* src = out + k * target; uint8_t* dst = good + k * req_comp;
            dst[0] = compute_y(src[0], src[1], src[2]);
            if (req_comp == 2) dst[1] = target == 4 ? src[3] : 255;
        }
        free(out); out = good;
    }
    *x = W; *y = H;
    if (comp) *comp = s->img_n;
    *ppmX = s->ppmX; *ppmY = s->ppmY; *pixelRatio = s->pixelAspectRatio;
    return out;
}

/* ---- format detection: Image.identifyFormatFromStream (image.d:1045-1061) over the plugins' detect procs
 * (plugins/jpeg.d:106, png.d:165, qoi.d:143, qoix.d:149, dds.d:40, gif.d:42, bmp.d:45, jxl.d:142, sqz.d:135, tga.d:97 +
 * TGADecoder.getImageInfo, codecs/tga.d:313-382). Returns the ImageFormat value (types.d:14-28) or -1. ---- */
static int starts(const uint8_t* d, size_t len, const void* sig, size_t n) { return len >= n && memcmp(d, sig, n) == 0; }
static int tga_info(const uint8_t* d, size_t len)
{
    size_t p = 0;
#define RD8(v) do { if (p >= len) return 0; (v) = d[p++]; } while (0)
#define RD16(v) do { if (p + 2 > len) return 0; (v) = d[p] | (d[p + 1] << 8); p += 2; } while (0)
#define SKIP(n) do { if (p + (n) > len) return 0; p += (n); } while (0)
    int dataOffset, cmapType, imageType, palStart, palLen, cmapSize, w, h, bpp;
    RD8(dataOffset); (void)dataOffset;
    RD8(cmapType); if (cmapType > 1) return 0;
    RD8(imageType);
    if (cmapType == 1) {
        if (imageType != 1 && imageType != 9) return 0;
        RD16(palStart); (void)palStart;
        RD16(palLen); if (palLen == 0) return 0;
        RD8(cmapSize);
        if (cmapSize != 8 && cmapSize != 15 && cmapSize != 16 && cmapSize != 24 && cmapSize != 32) return 0;
        SKIP(4);
    } else {
        if (imageType != 2 && imageType != 3 && imageType != 10 && imageType != 11) return 0;
        SKIP(9);
    }
    RD16(w); RD16(h);
    if (w < 1 || h < 1) return 0;
    RD8(bpp);
    if (cmapType == 1 && bpp != 8 && bpp != 16) return 0;
    if (bpp != 8 && bpp != 15 && bpp != 16 && bpp != 24 && bpp != 32) return 0;
    return 1;
#undef RD8
#undef RD16
#undef SKIP
}
int or_identify_format(const uint8_t* d, size_t len)
{
    if (starts(d, len, "\xff\xd8", 2)) return 0;                                  /* JPEG */
    if (starts(d, len, "\x89PNG\r\n\x1a\n", 8)) return 1;                         /* PNG */
    if (starts(d, len, "qoif", 4)) return 2;
    if (starts(d, len, "qoix", 4)) return 3;
    if (starts(d, len, "DDS ", 4)) return 4;
    /* TGA (5) is tried last */
    if (starts(d, len, "GIF87a", 6) || starts(d, len, "GIF89a", 6)) return 6;
    if (len >= 18 && d[0] == 'B' && d[1] == 'M') {                                /* detectBMP: 'BM', 12 bytes, header size */
        uint32_t ds = d[14] | (d[15] << 8) | (d[16] << 16) | ((uint32_t)d[17] << 24);
        if (ds == 12 || ds == 40 || ds == 52 || ds == 56 || ds == 108 || ds == 124) return 7;
    }
    if (starts(d, len, "\xff\x0a", 2)) return 8;                                  /* JXL */
    if (starts(d, len, "\xa5", 1)) return 9;                                      /* SQZ */
    if (tga_info(d, len)) return 5;
    return -1;
}

/* ------------------------------------------------------------------------------------------ */
/* saveBMP (plugins/bmp.d:166-194) -> write_bmp (codecs/bmpenc.d:25-113): rgb8 / rgba8 rows (type = PixelType value 9 / 12;
 * `data` = first scanline, pitchBytes signed) -> BMP file with a V4 header. ppmX / ppmY = Image.pixelsPerMeterX / Y, -1 when
 * unknown. The reference writes the row padding of 24-bit files from an uninitialised buffer (:40-43, :103); here it is 0. */
#include <math.h>
uint8_t* or_bmp_encode(const uint8_t* data, int type, int width, int height, int pitchBytes, float ppmX, float ppmY, int* out_len)
{
    const int chans = type == 9 ? 3 : type == 12 ? 4 : 0;
    if (!chans || width < 1 || height < 1 || width > 32767 || height > 32767) return NULL;
    enum { DIB_SIZE = 108 };
    const int linesize = width * chans, pad = 3 - ((linesize - 1) & 3);
    const int idat_offset = 14 + DIB_SIZE;
    const size_t filesize = (size_t)idat_offset + (size_t)height * (size_t)(linesize + pad);
    uint8_t* out = (uint8_t*)calloc(filesize, 1);
    if (!out) return NULL;
    uint8_t* hdr = out;
#define LE32(at, v) do { const uint32_t v_ = (uint32_t)(v); hdr[at] = (uint8_t)v_; hdr[(at) + 1] = (uint8_t)(v_ >> 8); hdr[(at) + 2] = (uint8_t)(v_ >> 16); hdr[(at) + 3] = (uint8_t)(v_ >> 24); } while (0)
    hdr[0] = 0x42; hdr[1] = 0x4d;
    LE32(2, filesize); LE32(10, idat_offset); LE32(14, DIB_SIZE); LE32(18, width); LE32(22, height);
    hdr[26] = 1; hdr[27] = 0; hdr[28] = (uint8_t)(chans * 8); hdr[29] = 0;
    LE32(30, chans == 3 ? 0 : 3);
    int ippmX = 0, ippmY = 0;
    if (ppmX != -1.0f) ippmX = (int)round(ppmX);
    if (ppmY != -1.0f) ippmY = (int)round(ppmY);
    LE32(38, ippmX); LE32(42, ippmY);
    if (chans == 4) { static const uint8_t b[16] = {0, 0, 0xff, 0, 0, 0xff, 0, 0, 0xff, 0, 0, 0, 0, 0, 0, 0xff}; memcpy(hdr + 54, b, 16); }
    hdr[70] = 'B'; hdr[71] = 'G'; hdr[72] = 'R'; hdr[73] = 's';
#undef LE32
    for (int y = 0; y < height; ++y) {
        const uint8_t* in = data + (ptrdiff_t)pitchBytes * (height - 1 - y);
        uint8_t* o = out + idat_offset + (size_t)y * (size_t)(linesize + pad);
        for (int x = 0; x < width; ++x) {
            o[chans * x] = in[chans * x + 2]; o[chans * x + 1] = in[chans * x + 1]; o[chans * x + 2] = in[chans * x];
            if (chans == 4) o[4 * x + 3] = in[4 * x + 3];
        }
    }
    *out_len = (int)filesize;
    return out;
}
