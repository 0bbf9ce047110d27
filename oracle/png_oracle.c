/*
 * png_oracle.c -- CPU restatement of the PNG path of source/gamut/codecs/stbdec.d
 * (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * Follows stbdec.d function by function (stb_image 2.27 port): context readers :780-893,
 * stbi__convert_format[16] :916-1200, 16<->8 :635-666, zlib wrapper :1267-1321,
 * stbi__create_png_image_raw :1406-1635, Adam7 :1637-1680, tRNS :1682-1730, palette :1732-1765,
 * stbi__parse_png_file :1777-2023, stbi__do_png :2025-2055, stbi__png_is16 :2090-2110.
 *
 * The inflate engine of the reference is the third-party `miniz` D package (dub.json:9,
 * ">=0.0.0 <2.0.0", source NOT in the reference tree). DEFLATE output is fully determined by
 * RFC 1950/1951, so system zlib is the engine here; the wrapper semantics of stbdec.d:1267-1321
 * are restated: zlib header checked, Adler-32 neither required nor verified (trusted_input),
 * trailing bytes tolerated, output buffer doubled (min 32 KiB) and decode restarted when it is too
 * small, failure when the guess exceeds 536,870,912.
 *
 * parity: PINNED by the reference's own fixtures (tests/golden: issue76.png KAT, buggy-miniz-chunk
 * length KAT, must-load set) and cross-checked against PIL.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

typedef uint8_t stbi_uc;
typedef uint16_t stbi__uint16;
typedef uint32_t stbi__uint32;

#define STBI_MAX_DIMENSIONS (1 << 24)

/* memory-stream context: what stbi__context + gamut MemoryFile callbacks reduce to */
typedef struct {
    const stbi_uc* buf; size_t len; size_t pos;
    stbi__uint32 img_x, img_y; int img_n, img_out_n;
    float ppmX, ppmY, pixelAspectRatio;
} ctx;

/* stbdec.d:794-803: past the end the reader yields 0 */
static stbi_uc get8(ctx* s) { if (s->pos < s->len) return s->buf[s->pos++]; return 0; }
static int at_eof(ctx* s) { return s->pos >= s->len; }                                   /* :805 */
static void skip(ctx* s, int n) { if (n == 0) return; if (n < 0) { s->pos = s->len; return; }
    if (s->len - s->pos < (size_t)n) s->pos = s->len; else s->pos += (size_t)n; }        /* :817 */
static int getn(ctx* s, stbi_uc* dst, int n) {                                            /* :841 */
    if (s->len - s->pos < (size_t)n) { size_t b = s->len - s->pos; memcpy(dst, s->buf + s->pos, b); s->pos = s->len; return 0; }
    memcpy(dst, s->buf + s->pos, (size_t)n); s->pos += (size_t)n; return 1; }
static int get16be(ctx* s) { int z = get8(s); return (z << 8) + get8(s); }                /* :867 */
static stbi__uint32 get32be(ctx* s) { stbi__uint32 z = (stbi__uint32)get16be(s); return (z << 16) + (stbi__uint32)get16be(s); } /* :873 */

static stbi_uc compute_y(int r, int g, int b) { return (stbi_uc)(((r * 77) + (g * 150) + (29 * b)) >> 8); }         /* :911 */
static stbi__uint16 compute_y_16(int r, int g, int b) { return (stbi__uint16)(((r * 77) + (g * 150) + (29 * b)) >> 8); } /* :1056 */

/* stbdec.d:916-1054 and :1061-1200 (same switch for both widths) */
#define CONVERT_FORMAT(NAME, T, MAXV, CY)                                                          \
static T* NAME(T* data, int img_n, int req_comp, unsigned x, unsigned y) {                         \
    if (req_comp == img_n) return data;                                                            \
    T* good = (T*)malloc((size_t)req_comp * x * y * sizeof(T));                                    \
    if (!good) { free(data); return NULL; }                                                        \
    for (int j = 0; j < (int)y; ++j) {                                                             \
        T* src = data + (size_t)j * x * img_n; T* dest = good + (size_t)j * x * req_comp; int i;   \
        switch (img_n * 8 + req_comp) {                                                            \
        case 1*8+2: for (i = x-1; i >= 0; --i, src += 1, dest += 2) { dest[0] = src[0]; dest[1] = MAXV; } break; \
        case 1*8+3: for (i = x-1; i >= 0; --i, src += 1, dest += 3) { dest[0] = dest[1] = dest[2] = src[0]; } break; \
        case 1*8+4: for (i = x-1; i >= 0; --i, src += 1, dest += 4) { dest[0] = dest[1] = dest[2] = src[0]; dest[3] = MAXV; } break; \
        case 2*8+1: for (i = x-1; i >= 0; --i, src += 2, dest += 1) { dest[0] = src[0]; } break;   \
        case 2*8+3: for (i = x-1; i >= 0; --i, src += 2, dest += 3) { dest[0] = dest[1] = dest[2] = src[0]; } break; \
        case 2*8+4: for (i = x-1; i >= 0; --i, src += 2, dest += 4) { dest[0] = dest[1] = dest[2] = src[0]; dest[3] = src[1]; } break; \
        case 3*8+4: for (i = x-1; i >= 0; --i, src += 3, dest += 4) { dest[0] = src[0]; dest[1] = src[1]; dest[2] = src[2]; dest[3] = MAXV; } break; \
        case 3*8+1: for (i = x-1; i >= 0; --i, src += 3, dest += 1) { dest[0] = CY(src[0], src[1], src[2]); } break; \
        case 3*8+2: for (i = x-1; i >= 0; --i, src += 3, dest += 2) { dest[0] = CY(src[0], src[1], src[2]); dest[1] = MAXV; } break; \
        case 4*8+1: for (i = x-1; i >= 0; --i, src += 4, dest += 1) { dest[0] = CY(src[0], src[1], src[2]); } break; \
        case 4*8+2: for (i = x-1; i >= 0; --i, src += 4, dest += 2) { dest[0] = CY(src[0], src[1], src[2]); dest[1] = src[3]; } break; \
        case 4*8+3: for (i = x-1; i >= 0; --i, src += 4, dest += 3) { dest[0] = src[0]; dest[1] = src[1]; dest[2] = src[2]; } break; \
        default: free(data); free(good); return NULL;                                              \
        } }                                                                                        \
    free(data); return good; }
CONVERT_FORMAT(convert_format,   stbi_uc,      255,    compute_y)
CONVERT_FORMAT(convert_format16, stbi__uint16, 0xffff, compute_y_16)

/* stbdec.d:635-649 */
static stbi_uc* convert_16_to_8(stbi__uint16* orig, int w, int h, int channels)
{
    int img_len = w * h * channels;
    stbi_uc* reduced = (stbi_uc*)malloc((size_t)img_len);
    if (!reduced) return NULL;
    for (int i = 0; i < img_len; ++i) reduced[i] = (stbi_uc)((orig[i] >> 8) & 0xFF);
    free(orig);
    return reduced;
}
/* stbdec.d:651-666 */
static stbi__uint16* convert_8_to_16(stbi_uc* orig, int w, int h, int channels)
{
    int img_len = w * h * channels;
    stbi__uint16* enlarged = (stbi__uint16*)malloc((size_t)img_len * 2);
    if (!enlarged) return NULL;
    for (int i = 0; i < img_len; ++i) enlarged[i] = (stbi__uint16)((orig[i] << 8) + orig[i]);
    free(orig);
    return enlarged;
}

/* One mz_uncompress3 attempt: returns 0 ok (sets *destLen), 1 = MZ_BUF_ERROR (output too small),
 * 2 = data error. parse_header: zlib header checked, then raw inflate (Adler-32 not read). */
static int uncompress_once(stbi_uc* out, size_t* destLen, const stbi_uc* in, size_t inLen, int parse_header)
{
    if (parse_header) {
        if (inLen < 2) return 2;
        unsigned cmf = in[0], flg = in[1];
        if (((cmf * 256 + flg) % 31 != 0) || (flg & 32) || ((cmf & 15) != 8)) return 2;
        in += 2; inLen -= 2;
    }
    z_stream zs; memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return 2;
    zs.next_in = (Bytef*)in; zs.avail_in = (uInt)inLen;
    zs.next_out = out; zs.avail_out = (uInt)*destLen;
    int r = inflate(&zs, Z_FINISH);
    int ret;
    if (r == Z_STREAM_END) { *destLen = zs.total_out; ret = 0; }
    else if ((r == Z_BUF_ERROR || r == Z_OK) && zs.avail_out == 0 && zs.avail_in != 0) ret = 1;
    else if ((r == Z_BUF_ERROR || r == Z_OK) && zs.avail_out == 0) {
        /* output full and input exhausted at the same time: miniz reports a data error
         * (status==MZ_BUF_ERROR && !avail_in => MZ_DATA_ERROR) */
        ret = 2;
    }
    else ret = 2;
    inflateEnd(&zs);
    return ret;
}

/* stbdec.d:1267-1321 */
uint8_t* or_zlib_decode(const uint8_t* buffer, size_t len, size_t initial_size, int parse_header, size_t* outlen)
{
    stbi_uc* outBuf = (stbi_uc*)malloc(initial_size ? initial_size : 1);
    if (!outBuf) return NULL;
    size_t destLen = initial_size;
    for (;;) {
        int res = uncompress_once(outBuf, &destLen, buffer, len, parse_header);
        if (res == 0) break;
        if (res == 1) {
            if (initial_size > 536870912u) { free(outBuf); return NULL; }
            initial_size = initial_size * 2;
            if (initial_size < 32 * 1024) initial_size = 32 * 1024;
            outBuf = (stbi_uc*)realloc(outBuf, initial_size);
            if (!outBuf) return NULL;
            destLen = initial_size;
        } else { free(outBuf); return NULL; }
    }
    *outlen = destLen;
    return outBuf;
}

typedef struct { ctx* s; stbi_uc* idata; stbi_uc* expanded; stbi_uc* out_; int depth; } png;

enum { F_none = 0, F_sub = 1, F_up = 2, F_avg = 3, F_paeth = 4, F_avg_first, F_paeth_first };
static const stbi_uc first_row_filter[5] = { F_none, F_sub, F_none, F_avg_first, F_paeth_first };   /* :1381 */

static int paeth(int a, int b, int c)                                                               /* :1390 */
{
    int p = a + b - c;
    int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    if (pb <= pc) return b;
    return c;
}
static const stbi_uc depth_scale_table[9] = { 0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01 };              /* :1403 */
#define BYTECAST(x) ((stbi_uc)((x) & 255))

/* stbdec.d:1406-1635 */
static int create_png_image_raw(png* a, stbi_uc* raw, stbi__uint32 raw_len, int out_n,
                                stbi__uint32 x, stbi__uint32 y, int depth, int color)
{
    int bytes = (depth == 16 ? 2 : 1);
    ctx* s = a->s;
    stbi__uint32 i, j, stride = x * out_n * bytes;
    stbi__uint32 img_len, img_width_bytes;
    int k;
    int img_n = s->img_n;
    int output_bytes = out_n * bytes;
    int filter_bytes = img_n * bytes;
    int width = (int)x;

    a->out_ = (stbi_uc*)malloc((size_t)x * y * output_bytes + 16);
    if (!a->out_) return 0;
    img_width_bytes = (((img_n * x * depth) + 7) >> 3);
    img_len = (img_width_bytes + 1) * y;
    if (raw_len < img_len) return 0;

    for (j = 0; j < y; ++j) {
        stbi_uc* cur = a->out_ + (size_t)stride * j;
        stbi_uc* prior;
        int filter = *raw++;
        if (filter > 4) return 0;
        if (depth < 8) {
            if (img_width_bytes > x) return 0;
            cur += x * out_n - img_width_bytes;
            filter_bytes = 1;
            width = (int)img_width_bytes;
        }
        prior = cur - stride;
        if (j == 0) filter = first_row_filter[filter];

        for (k = 0; k < filter_bytes; ++k) {
            switch (filter) {
            case F_none: cur[k] = raw[k]; break;
            case F_sub: cur[k] = raw[k]; break;
            case F_up: cur[k] = BYTECAST(raw[k] + prior[k]); break;
            case F_avg: cur[k] = BYTECAST(raw[k] + (prior[k] >> 1)); break;
            case F_paeth: cur[k] = BYTECAST(raw[k] + paeth(0, prior[k], 0)); break;
            case F_avg_first: cur[k] = raw[k]; break;
            case F_paeth_first: cur[k] = raw[k]; break;
            }
        }
        if (depth == 8) {
            if (img_n != out_n) cur[img_n] = 255;
            raw += img_n; cur += out_n; prior += out_n;
        } else if (depth == 16) {
            if (img_n != out_n) { cur[filter_bytes] = 255; cur[filter_bytes + 1] = 255; }
            raw += filter_bytes; cur += output_bytes; prior += output_bytes;
        } else { raw += 1; cur += 1; prior += 1; }

        if (depth < 8 || img_n == out_n) {
            int nk = (width - 1) * filter_bytes;
            switch (filter) {
            case F_none: memcpy(cur, raw, (size_t)nk); break;
            case F_sub: for (k = 0; k < nk; ++k) cur[k] = BYTECAST(raw[k] + cur[k - filter_bytes]); break;
            case F_up: for (k = 0; k < nk; ++k) cur[k] = BYTECAST(raw[k] + prior[k]); break;
            case F_avg: for (k = 0; k < nk; ++k) cur[k] = BYTECAST(raw[k] + ((prior[k] + cur[k - filter_bytes]) >> 1)); break;
            case F_paeth: for (k = 0; k < nk; ++k) cur[k] = BYTECAST(raw[k] + paeth(cur[k - filter_bytes], prior[k], prior[k - filter_bytes])); break;
            case F_avg_first: for (k = 0; k < nk; ++k) cur[k] = BYTECAST(raw[k] + (cur[k - filter_bytes] >> 1)); break;
            case F_paeth_first: for (k = 0; k < nk; ++k) cur[k] = BYTECAST(raw[k] + paeth(cur[k - filter_bytes], 0, 0)); break;
            }
            raw += nk;
        } else {
#define ROWLOOP(EXPR) for (i = x - 1; i >= 1; --i, cur[filter_bytes] = 255, raw += filter_bytes, cur += output_bytes, prior += output_bytes) \
                          for (k = 0; k < filter_bytes; ++k) { cur[k] = (EXPR); }
            switch (filter) {
            case F_none: ROWLOOP(raw[k]) break;
            case F_sub: ROWLOOP(BYTECAST(raw[k] + cur[k - output_bytes])) break;
            case F_up: ROWLOOP(BYTECAST(raw[k] + prior[k])) break;
            case F_avg: ROWLOOP(BYTECAST(raw[k] + ((prior[k] + cur[k - output_bytes]) >> 1))) break;
            case F_paeth: ROWLOOP(BYTECAST(raw[k] + paeth(cur[k - output_bytes], prior[k], prior[k - output_bytes]))) break;
            case F_avg_first: ROWLOOP(BYTECAST(raw[k] + (cur[k - output_bytes] >> 1))) break;
            case F_paeth_first: ROWLOOP(BYTECAST(raw[k] + paeth(cur[k - output_bytes], 0, 0))) break;
            }
#undef ROWLOOP
            if (depth == 16) {
                cur = a->out_ + (size_t)stride * j;
                for (i = 0; i < x; ++i, cur += output_bytes) cur[filter_bytes + 1] = 255;
            }
        }
    }

    if (depth < 8) {
        for (j = 0; j < y; ++j) {
            stbi_uc* cur = a->out_ + (size_t)stride * j;
            stbi_uc* in_ = a->out_ + (size_t)stride * j + x * out_n - img_width_bytes;
            stbi_uc scale = (color == 0) ? depth_scale_table[depth] : 1;
            if (depth == 4) {
                for (k = x * img_n; k >= 2; k -= 2, ++in_) {
                    *cur++ = (stbi_uc)(scale * ((*in_ >> 4)));
                    *cur++ = (stbi_uc)(scale * ((*in_) & 0x0f));
                }
                if (k > 0) *cur++ = (stbi_uc)(scale * ((*in_ >> 4)));
            } else if (depth == 2) {
                for (k = x * img_n; k >= 4; k -= 4, ++in_) {
                    *cur++ = (stbi_uc)(scale * ((*in_ >> 6)));
                    *cur++ = (stbi_uc)(scale * ((*in_ >> 4) & 0x03));
                    *cur++ = (stbi_uc)(scale * ((*in_ >> 2) & 0x03));
                    *cur++ = (stbi_uc)(scale * ((*in_) & 0x03));
                }
                if (k > 0) *cur++ = (stbi_uc)(scale * ((*in_ >> 6)));
                if (k > 1) *cur++ = (stbi_uc)(scale * ((*in_ >> 4) & 0x03));
                if (k > 2) *cur++ = (stbi_uc)(scale * ((*in_ >> 2) & 0x03));
            } else if (depth == 1) {
                for (k = x * img_n; k >= 8; k -= 8, ++in_) {
                    *cur++ = (stbi_uc)(scale * ((*in_ >> 7)));
                    *cur++ = (stbi_uc)(scale * ((*in_ >> 6) & 0x01));
                    *cur++ = (stbi_uc)(scale * ((*in_ >> 5) & 0x01));
                    *cur++ = (stbi_uc)(scale * ((*in_ >> 4) & 0x01));
                    *cur++ = (stbi_uc)(scale * ((*in_ >> 3) & 0x01));
                    *cur++ = (stbi_uc)(scale * ((*in_ >> 2) & 0x01));
                    *cur++ = (stbi_uc)(scale * ((*in_ >> 1) & 0x01));
                    *cur++ = (stbi_uc)(scale * ((*in_) & 0x01));
                }
                if (k > 0) *cur++ = (stbi_uc)(scale * ((*in_ >> 7)));
                if (k > 1) *cur++ = (stbi_uc)(scale * ((*in_ >> 6) & 0x01));
                if (k > 2) *cur++ = (stbi_uc)(scale * ((*in_ >> 5) & 0x01));
                if (k > 3) *cur++ = (stbi_uc)(scale * ((*in_ >> 4) & 0x01));
                if (k > 4) *cur++ = (stbi_uc)(scale * ((*in_ >> 3) & 0x01));
                if (k > 5) *cur++ = (stbi_uc)(scale * ((*in_ >> 2) & 0x01));
                if (k > 6) *cur++ = (stbi_uc)(scale * ((*in_ >> 1) & 0x01));
            }
            if (img_n != out_n) {
                int q;
                cur = a->out_ + (size_t)stride * j;
                if (img_n == 1) {
                    for (q = (int)x - 1; q >= 0; --q) { cur[q*2+1] = 255; cur[q*2+0] = cur[q]; }
                } else {
                    for (q = (int)x - 1; q >= 0; --q) {
                        cur[q*4+3] = 255; cur[q*4+2] = cur[q*3+2]; cur[q*4+1] = cur[q*3+1]; cur[q*4+0] = cur[q*3+0];
                    }
                }
            }
        }
    } else if (depth == 16) {
        stbi_uc* cur = a->out_;
        stbi__uint16* cur16 = (stbi__uint16*)cur;
        for (i = 0; i < x * y * out_n; ++i, cur16++, cur += 2) *cur16 = (stbi__uint16)((cur[0] << 8) | cur[1]);
    }
    return 1;
}

/* stbdec.d:1637-1680 */
static int create_png_image(png* a, stbi_uc* image_data, stbi__uint32 image_data_len, int out_n, int depth, int color, int interlaced)
{
    int bytes = (depth == 16 ? 2 : 1);
    int out_bytes = out_n * bytes;
    stbi_uc* final_;
    int p;
    if (!interlaced)
        return create_png_image_raw(a, image_data, image_data_len, out_n, a->s->img_x, a->s->img_y, depth, color);

    final_ = (stbi_uc*)malloc((size_t)a->s->img_x * a->s->img_y * out_bytes);
    if (!final_) return 0;
    for (p = 0; p < 7; ++p) {
        static const int xorig[7] = { 0,4,0,2,0,1,0 };
        static const int yorig[7] = { 0,0,4,0,2,0,1 };
        static const int xspc[7]  = { 8,8,4,4,2,2,1 };
        static const int yspc[7]  = { 8,8,8,4,4,2,2 };
        int i, j, x, y;
        x = (int)((a->s->img_x - xorig[p] + xspc[p] - 1) / xspc[p]);
        y = (int)((a->s->img_y - yorig[p] + yspc[p] - 1) / yspc[p]);
        if (x && y) {
            stbi__uint32 img_len = ((((a->s->img_n * x * depth) + 7) >> 3) + 1) * y;
            if (!create_png_image_raw(a, image_data, image_data_len, out_n, x, y, depth, color)) {
                free(final_); free(a->out_); a->out_ = NULL;
                return 0;
            }
            for (j = 0; j < y; ++j) {
                for (i = 0; i < x; ++i) {
                    int out_y = j * yspc[p] + yorig[p];
                    int out_x = i * xspc[p] + xorig[p];
                    memcpy(final_ + (size_t)out_y * a->s->img_x * out_bytes + (size_t)out_x * out_bytes,
                           a->out_ + ((size_t)j * x + i) * out_bytes, (size_t)out_bytes);
                }
            }
            free(a->out_);
            image_data += img_len;
            image_data_len -= img_len;
        }
    }
    a->out_ = final_;
    return 1;
}

/* stbdec.d:1682-1730 */
static void compute_transparency(png* z, stbi_uc* tc, int out_n)
{
    ctx* s = z->s; stbi__uint32 i, pixel_count = s->img_x * s->img_y; stbi_uc* p = z->out_;
    if (out_n == 2) { for (i = 0; i < pixel_count; ++i) { p[1] = (p[0] == tc[0] ? 0 : 255); p += 2; } }
    else { for (i = 0; i < pixel_count; ++i) { if (p[0] == tc[0] && p[1] == tc[1] && p[2] == tc[2]) p[3] = 0; p += 4; } }
}
static void compute_transparency16(png* z, stbi__uint16* tc, int out_n)
{
    ctx* s = z->s; stbi__uint32 i, pixel_count = s->img_x * s->img_y; stbi__uint16* p = (stbi__uint16*)z->out_;
    if (out_n == 2) { for (i = 0; i < pixel_count; ++i) { p[1] = (p[0] == tc[0] ? 0 : 65535); p += 2; } }
    else { for (i = 0; i < pixel_count; ++i) { if (p[0] == tc[0] && p[1] == tc[1] && p[2] == tc[2]) p[3] = 0; p += 4; } }
}

/* stbdec.d:1732-1765 */
static int expand_png_palette(png* a, stbi_uc* palette, int len, int pal_img_n)
{
    stbi__uint32 i, pixel_count = a->s->img_x * a->s->img_y;
    stbi_uc *p, *temp_out, *orig = a->out_;
    (void)len;
    p = (stbi_uc*)malloc((size_t)pixel_count * pal_img_n);
    if (!p) return 0;
    temp_out = p;
    if (pal_img_n == 3) {
        for (i = 0; i < pixel_count; ++i) { int n = orig[i] * 4; p[0] = palette[n]; p[1] = palette[n+1]; p[2] = palette[n+2]; p += 3; }
    } else {
        for (i = 0; i < pixel_count; ++i) { int n = orig[i] * 4; p[0] = palette[n]; p[1] = palette[n+1]; p[2] = palette[n+2]; p[3] = palette[n+3]; p += 4; }
    }
    free(a->out_);
    a->out_ = temp_out;
    return 1;
}

#define PNG_TYPE(a,b,c,d) (((unsigned)(a) << 24) + ((unsigned)(b) << 16) + ((unsigned)(c) << 8) + (unsigned)(d))
enum { SCAN_load = 0, SCAN_type, SCAN_header };

typedef struct {
    stbi_uc palette[1024]; stbi_uc pal_img_n, has_trans; stbi_uc tc[3]; stbi__uint16 tc16[3];
    stbi__uint32 ioff, idata_limit, pal_len; int interlace, color, is_iphone;
} pstate;

/* nested finalize_decode, stbdec.d:1800-1858 */
static int finalize_decode(png* z, pstate* P, int scan, int req_comp)
{
    ctx* s = z->s;
    stbi__uint32 raw_len, bpl;
    if (scan != SCAN_load) return 1;
    if (z->idata == NULL) return 0;
    bpl = (s->img_x * z->depth + 7) / 8;
    raw_len = bpl * s->img_y * s->img_n + s->img_y;
    size_t outlen = 0;
    z->expanded = or_zlib_decode(z->idata, P->ioff, raw_len, !P->is_iphone, &outlen);
    if (z->expanded == NULL) return 0;
    raw_len = (stbi__uint32)(int)outlen;
    free(z->idata); z->idata = NULL;
    if ((req_comp == s->img_n + 1 && req_comp != 3 && !P->pal_img_n) || P->has_trans)
        s->img_out_n = s->img_n + 1;
    else
        s->img_out_n = s->img_n;
    if (!create_png_image(z, z->expanded, raw_len, s->img_out_n, z->depth, P->color, P->interlace)) return 0;
    if (P->has_trans) {
        if (z->depth == 16) compute_transparency16(z, P->tc16, s->img_out_n);
        else compute_transparency(z, P->tc, s->img_out_n);
    }
    if (P->pal_img_n) {
        s->img_n = P->pal_img_n;
        s->img_out_n = P->pal_img_n;
        if (req_comp >= 3) s->img_out_n = req_comp;
        if (!expand_png_palette(z, P->palette, (int)P->pal_len, s->img_out_n)) return 0;
    } else if (P->has_trans) {
        ++s->img_n;
    }
    free(z->expanded); z->expanded = NULL;
    return 1;
}

/* stbdec.d:1777-2023 */
static int parse_png_file(png* z, int scan, int req_comp)
{
    pstate P; memset(&P, 0, sizeof(P));   /* NOTE: the reference leaves palette[] uninitialised */
    stbi__uint32 i; int first = 1, k;
    ctx* s = z->s;
    static const stbi_uc png_sig[8] = { 137,80,78,71,13,10,26,10 };

    z->expanded = NULL; z->idata = NULL; z->out_ = NULL;
    s->ppmX = -1; s->ppmY = -1; s->pixelAspectRatio = -1;

    for (i = 0; i < 8; ++i) if (get8(s) != png_sig[i]) return 0;     /* stbi__check_png_header :1349 */
    if (scan == SCAN_type) return 1;

    for (;;) {
        stbi__uint32 c_length = get32be(s);
        stbi__uint32 c_type = get32be(s);
        switch (c_type) {
        case PNG_TYPE('C','g','B','I'):
            P.is_iphone = 1; skip(s, (int)c_length); break;
        case PNG_TYPE('p','H','Y','s'): {
            s->ppmX = (float)get32be(s);
            s->ppmY = (float)get32be(s);
            s->pixelAspectRatio = s->ppmX / s->ppmY;
            stbi_uc unit = get8(s);
            if (unit != 1) { s->ppmX = -1; s->ppmY = -1; }
            break; }
        case PNG_TYPE('I','H','D','R'): {
            int comp, filter;
            if (!first) return 0;
            first = 0;
            if (c_length != 13) return 0;
            s->img_x = get32be(s);
            s->img_y = get32be(s);
            if (s->img_y > STBI_MAX_DIMENSIONS) return 0;
            if (s->img_x > STBI_MAX_DIMENSIONS) return 0;
            z->depth = get8(s);
            if (z->depth != 1 && z->depth != 2 && z->depth != 4 && z->depth != 8 && z->depth != 16) return 0;
            P.color = get8(s); if (P.color > 6) return 0;
            if (P.color == 3 && z->depth == 16) return 0;
            if (P.color == 3) P.pal_img_n = 3; else if (P.color & 1) return 0;
            comp = get8(s); if (comp) return 0;
            filter = get8(s); if (filter) return 0;
            P.interlace = get8(s); if (P.interlace > 1) return 0;
            if (!s->img_x || !s->img_y) return 0;
            if (!P.pal_img_n) {
                s->img_n = (P.color & 2 ? 3 : 1) + (P.color & 4 ? 1 : 0);
                if ((1 << 30) / s->img_x / s->img_n < s->img_y) return 0;
                if (scan == SCAN_header) return 1;
            } else {
                s->img_n = 1;
                if ((1 << 30) / s->img_x / 4 < s->img_y) return 0;
            }
            break; }
        case PNG_TYPE('P','L','T','E'): {
            if (first) return 0;
            if (c_length > 256 * 3) return 0;
            P.pal_len = c_length / 3;
            if (P.pal_len * 3 != c_length) return 0;
            for (i = 0; i < P.pal_len; ++i) {
                P.palette[i*4+0] = get8(s); P.palette[i*4+1] = get8(s); P.palette[i*4+2] = get8(s); P.palette[i*4+3] = 255;
            }
            break; }
        case PNG_TYPE('t','R','N','S'): {
            if (first) return 0;
            if (z->idata) return 0;
            if (P.pal_img_n) {
                if (scan == SCAN_header) { s->img_n = 4; return 1; }
                if (P.pal_len == 0) return 0;
                if (c_length > P.pal_len) return 0;
                P.pal_img_n = 4;
                for (i = 0; i < c_length; ++i) P.palette[i*4+3] = get8(s);
            } else {
                if (!(s->img_n & 1)) return 0;
                if (c_length != (stbi__uint32)s->img_n * 2) return 0;
                P.has_trans = 1;
                if (z->depth == 16) {
                    for (k = 0; k < s->img_n; ++k) P.tc16[k] = (stbi__uint16)get16be(s);
                } else {
                    for (k = 0; k < s->img_n; ++k)
                        P.tc[k] = (stbi_uc)((stbi_uc)(get16be(s) & 255) * depth_scale_table[z->depth]);
                }
            }
            break; }
        case PNG_TYPE('I','D','A','T'): {
            if (first) return 0;
            if (P.pal_img_n && !P.pal_len) return 0;
            if (scan == SCAN_header) { s->img_n = P.pal_img_n; return 1; }
            if ((int)(P.ioff + c_length) < (int)P.ioff) return 0;
            if (P.ioff + c_length > P.idata_limit) {
                stbi_uc* p;
                if (P.idata_limit == 0) P.idata_limit = c_length > 4096 ? c_length : 4096;
                while (P.ioff + c_length > P.idata_limit) P.idata_limit *= 2;
                p = (stbi_uc*)realloc(z->idata, P.idata_limit);
                if (p == NULL) return 0;
                z->idata = p;
            }
            if (!getn(s, z->idata + P.ioff, (int)c_length)) return 0;
            P.ioff += c_length;
            break; }
        case PNG_TYPE('I','E','N','D'): {
            if (first) return 0;
            int res = finalize_decode(z, &P, scan, req_comp);
            if (!res) return res;
            get32be(s);
            return 1; }
        default:
            if (first) return 0;
            if (c_type == 0 && at_eof(s)) return finalize_decode(z, &P, scan, req_comp);   /* gamut issue #92 */
            if ((c_type & (1 << 29)) == 0) return 0;
            skip(s, (int)c_length);
            break;
        }
        get32be(s);   /* CRC, not checked */
    }
}

/* stbdec.d:2025-2055 (+ stbi__png_load :2057) */
static void* do_png(png* p, int* x, int* y, int* n, int req_comp, int* bits_per_channel)
{
    void* result = NULL;
    if (req_comp < 0 || req_comp > 4) return NULL;
    if (parse_png_file(p, SCAN_load, req_comp)) {
        if (p->depth <= 8) *bits_per_channel = 8;
        else if (p->depth == 16) *bits_per_channel = 16;
        else return NULL;
        result = p->out_;
        p->out_ = NULL;
        if (req_comp && req_comp != p->s->img_out_n) {
            if (*bits_per_channel == 8)
                result = convert_format((stbi_uc*)result, p->s->img_out_n, req_comp, p->s->img_x, p->s->img_y);
            else
                result = convert_format16((stbi__uint16*)result, p->s->img_out_n, req_comp, p->s->img_x, p->s->img_y);
            p->s->img_out_n = req_comp;
            if (result == NULL) return result;
        }
        *x = (int)p->s->img_x; *y = (int)p->s->img_y;
        if (n) *n = p->s->img_n;
    }
    free(p->out_); p->out_ = NULL;
    free(p->expanded); p->expanded = NULL;
    free(p->idata); p->idata = NULL;
    return result;
}

/* stbi__png_is16 (stbdec.d:2090-2110) */
int or_png_is16(const uint8_t* data, size_t len)
{
    ctx s; memset(&s, 0, sizeof(s)); s.buf = data; s.len = len;
    png p; memset(&p, 0, sizeof(p)); p.s = &s;
    if (!parse_png_file(&p, SCAN_header, 0)) return 0;
    return p.depth == 16;
}

/* stbi_load_from_callbacks / stbi_load_16_from_callbacks (stbdec.d:713-735) with
 * stbi__load_and_postprocess_8bit/16bit (:669-709) */
uint8_t* or_png_load(const uint8_t* data, size_t len, int req_comp, int want16, or_png_info* info)
{
    ctx s; memset(&s, 0, sizeof(s)); s.buf = data; s.len = len;
    png p; memset(&p, 0, sizeof(p)); p.s = &s;
    int x = 0, y = 0, comp = 0, bpc = 8;
    /* stbi__png_test :2064 */
    static const stbi_uc png_sig[8] = { 137,80,78,71,13,10,26,10 };
    if (len < 8 || memcmp(data, png_sig, 8) != 0) return NULL;
    void* result = do_png(&p, &x, &y, &comp, req_comp, &bpc);
    if (info) { info->ppmX = s.ppmX; info->ppmY = s.ppmY; info->pixelRatio = s.pixelAspectRatio; }
    if (!result) return NULL;
    int ch = req_comp == 0 ? comp : req_comp;
    if (!want16 && bpc != 8) result = convert_16_to_8((stbi__uint16*)result, x, y, ch);
    else if (want16 && bpc != 16) result = convert_8_to_16((stbi_uc*)result, x, y, ch);
    if (info) {
        info->width = x; info->height = y; info->file_channels = comp; info->channels = ch;
        info->bits = want16 ? 16 : 8;
    }
    return (uint8_t*)result;
}

/* Unfilter-only entry (stbi__create_png_image_raw) for kernel-level parity tests. */
int or_png_unfilter(const uint8_t* raw, size_t raw_len, int img_n, int out_n, int w, int h, int depth, uint8_t* out)
{
    ctx s; memset(&s, 0, sizeof(s)); s.img_n = img_n; s.img_x = (stbi__uint32)w; s.img_y = (stbi__uint32)h;
    png p; memset(&p, 0, sizeof(p)); p.s = &s; p.depth = depth;
    int ok = create_png_image_raw(&p, (stbi_uc*)raw, (stbi__uint32)raw_len, out_n, (stbi__uint32)w, (stbi__uint32)h, depth, 0);
    if (ok) memcpy(out, p.out_, (size_t)w * h * out_n * (depth == 16 ? 2 : 1));
    free(p.out_);
    return ok;
}
