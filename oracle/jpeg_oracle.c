/*
 * jpeg_oracle.c -- CPU restatement of the baseline (sequential Huffman) JPEG path of
 * source/gamut/codecs/jpegload.d (jpgd port). TEST INFRASTRUCTURE ONLY; see oracle.h.
 *
 * Follows jpegload.d: constants :106-145, Row/Col IDCT :156-292, idct :308-376, idct_4x4 :378-397,
 * bit reader :640-743, huff_decode :746-813 (restated as canonical decode, see below),
 * HUFF_EXTEND :816-822, DCT_Upsample :827-1073, markers :1177-1543, process_markers :1578-1845,
 * locate_* :1851-1967, create_look_ups :2080-2094, transform_mcu[_expand] :2120-2255,
 * process_restart :2335-2402, decode_next_row :2405-2525, H*Convert/gray/expanded :2528-2823,
 * make_huff_table :2851-2987, calc_mcu_block_order :3038-3090, init_frame :3130-3268,
 * decompress_jpeg_image_from_stream :3720-3808.
 *
 * Deliberate restatement choices (all equivalent on valid streams):
 *  - the whole file is in memory; past its end the byte source yields FF D9 FF D9 ... (:640-655);
 *  - Huffman decode is canonical (JPEG Annex C codes, which is what make_huff_table builds); a bit
 *    pattern that is not a code word is a decode failure here, where the reference walks an unset
 *    tree entry (garbage on corrupt streams only);
 *  - the sparse Row!N/Col!N IDCT instantiations and P_Q!(R,C)/R_S!(R,C) are evaluated densely on the
 *    zero-filled block (algebraically identical integer expressions);
 *  - progressive (SOF2) streams: restated in round 2 (init_progressive, decode_scan, load_next_row);
 *  - find_eoi (:2826-2848) is not restated (it cannot change pixels of a successful decode).
 *
 * parity: pixel values are UNPINNED by the reference (its only JPEG assertion is "issue35.jpg loads,
 * issue46.jpg fails"); cross-checked against libjpeg-turbo ISLOW (PIL) within +-1 for 4:4:4/grey in
 * tests/test_oracle_jpeg.py.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

typedef int16_t jpgd_block_t;
typedef int16_t jpgd_quant_t;

enum { MAX_BLOCKS_PER_MCU = 10, MAX_HUFF_TABLES = 8, MAX_QUANT_TABLES = 4, MAX_COMPONENTS = 4, MAX_COMPS_IN_SCAN = 4,
       MAX_BLOCKS_PER_ROW = 8192, MAX_HEIGHT = 16384, MAX_WIDTH = 16384 };
enum { GRAYSCALE = 0, YH1V1, YH2V1, YH1V2, YH2V2 };
enum { M_SOF0 = 0xC0, M_SOF1 = 0xC1, M_SOF2 = 0xC2, M_SOF3 = 0xC3, M_SOF5 = 0xC5, M_SOF6 = 0xC6, M_SOF7 = 0xC7, M_JPG = 0xC8,
       M_SOF9 = 0xC9, M_SOF10 = 0xCA, M_SOF11 = 0xCB, M_SOF13 = 0xCD, M_SOF14 = 0xCE, M_SOF15 = 0xCF, M_DHT = 0xC4, M_DAC = 0xCC,
       M_RST0 = 0xD0, M_RST7 = 0xD7, M_SOI = 0xD8, M_EOI = 0xD9, M_SOS = 0xDA, M_DQT = 0xDB, M_DRI = 0xDD, M_APP0 = 0xE0, M_TEM = 0x01 };

static const int g_ZAG[64] = { 0,1,8,16,9,2,3,10,17,24,32,25,18,11,4,5,12,19,26,33,40,48,41,34,27,20,13,6,7,14,21,28,35,42,49,56,57,50,43,36,29,22,15,23,30,37,44,51,58,59,52,45,38,31,39,46,53,60,61,54,47,55,62,63 };

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
static inline int DESCALE(int x, int n) { return (x + (1 << (n - 1))) >> n; }
static inline int DESCALE_ZEROSHIFT(int x, int n) { return (x + (128 << n) + (1 << (n - 1))) >> n; }
static inline uint8_t CLAMP(int i) { if (i < 0) i = 0; if (i > 255) i = 255; return (uint8_t)i; }

/* Row!(8).idct with zeros where the block has none (jpegload.d:156-214); ncols limits the columns read */
static void row_idct(int* pTemp, const jpgd_block_t* pSrc, int ncols)
{
#define AC(x) ((x) < ncols ? (int)pSrc[x] : 0)
    const int z2 = AC(2), z3 = AC(6);
    const int z1 = (z2 + z3) * FIX_0_541196100;
    const int tmp2 = z1 + z3 * (-FIX_1_847759065);
    const int tmp3 = z1 + z2 * FIX_0_765366865;
    const int tmp0 = (int)((unsigned)(AC(0) + AC(4)) << CONST_BITS);
    const int tmp1 = (int)((unsigned)(AC(0) - AC(4)) << CONST_BITS);
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    const int atmp0 = AC(7), atmp1 = AC(5), atmp2 = AC(3), atmp3 = AC(1);
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
    pTemp[0] = DESCALE(tmp10 + btmp3, CONST_BITS - PASS1_BITS);
    pTemp[7] = DESCALE(tmp10 - btmp3, CONST_BITS - PASS1_BITS);
    pTemp[1] = DESCALE(tmp11 + btmp2, CONST_BITS - PASS1_BITS);
    pTemp[6] = DESCALE(tmp11 - btmp2, CONST_BITS - PASS1_BITS);
    pTemp[2] = DESCALE(tmp12 + btmp1, CONST_BITS - PASS1_BITS);
    pTemp[5] = DESCALE(tmp12 - btmp1, CONST_BITS - PASS1_BITS);
    pTemp[3] = DESCALE(tmp13 + btmp0, CONST_BITS - PASS1_BITS);
    pTemp[4] = DESCALE(tmp13 - btmp0, CONST_BITS - PASS1_BITS);
#undef AC
}

/* Col!(8).idct (jpegload.d:218-292); nrows limits the rows read */
static void col_idct(uint8_t* pDst_ptr, const int* pTemp, int nrows)
{
#define AR(x) ((x) < nrows ? pTemp[(x) * 8] : 0)
    const int z2 = AR(2), z3 = AR(6);
    const int z1 = (z2 + z3) * FIX_0_541196100;
    const int tmp2 = z1 + z3 * (-FIX_1_847759065);
    const int tmp3 = z1 + z2 * FIX_0_765366865;
    const int tmp0 = (int)((unsigned)(AR(0) + AR(4)) << CONST_BITS);
    const int tmp1 = (int)((unsigned)(AR(0) - AR(4)) << CONST_BITS);
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    const int atmp0 = AR(7), atmp1 = AR(5), atmp2 = AR(3), atmp3 = AR(1);
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
    pDst_ptr[8*0] = CLAMP(DESCALE_ZEROSHIFT(tmp10 + btmp3, CONST_BITS + PASS1_BITS + 3));
    pDst_ptr[8*7] = CLAMP(DESCALE_ZEROSHIFT(tmp10 - btmp3, CONST_BITS + PASS1_BITS + 3));
    pDst_ptr[8*1] = CLAMP(DESCALE_ZEROSHIFT(tmp11 + btmp2, CONST_BITS + PASS1_BITS + 3));
    pDst_ptr[8*6] = CLAMP(DESCALE_ZEROSHIFT(tmp11 - btmp2, CONST_BITS + PASS1_BITS + 3));
    pDst_ptr[8*2] = CLAMP(DESCALE_ZEROSHIFT(tmp12 + btmp1, CONST_BITS + PASS1_BITS + 3));
    pDst_ptr[8*5] = CLAMP(DESCALE_ZEROSHIFT(tmp12 - btmp1, CONST_BITS + PASS1_BITS + 3));
    pDst_ptr[8*3] = CLAMP(DESCALE_ZEROSHIFT(tmp13 + btmp0, CONST_BITS + PASS1_BITS + 3));
    pDst_ptr[8*4] = CLAMP(DESCALE_ZEROSHIFT(tmp13 - btmp0, CONST_BITS + PASS1_BITS + 3));
#undef AR
}

/* idct (jpegload.d:308-376): dense evaluation; the block is zero beyond block_max_zag */
static void idct(const jpgd_block_t* pSrc_ptr, uint8_t* pDst_ptr, int block_max_zag)
{
    if (block_max_zag <= 1) {
        int k = ((pSrc_ptr[0] + 4) >> 3) + 128;
        k = CLAMP(k);
        memset(pDst_ptr, k, 64);
        return;
    }
    int temp[64];
    for (int i = 0; i < 8; ++i) row_idct(temp + i * 8, pSrc_ptr + i * 8, 8);
    for (int i = 0; i < 8; ++i) col_idct(pDst_ptr + i, temp + i, 8);
}

/* idct_4x4 (jpegload.d:378-397): Row!4 on rows 0..3, Col!4 */
static void idct_4x4(const jpgd_block_t* pSrc_ptr, uint8_t* pDst_ptr)
{
    int temp[64];
    memset(temp, 0, sizeof(temp));
    for (int i = 0; i < 4; ++i) row_idct(temp + i * 8, pSrc_ptr + i * 8, 4);
    for (int i = 0; i < 8; ++i) col_idct(pDst_ptr + i, temp + i, 4);
}

/* ---- DCT_Upsample (jpegload.d:827-1073) ---- */
#define UF(x) ((int)((x) * 1024 + 0.5f))      /* F!(x), FRACT_BITS = 10 */
static inline int UD(int i) { return (i + 512) >> 10; }
typedef struct { int v[4][4]; } Matrix44;

/* P_Q!(8,8) and R_S!(8,8) on the zero-filled block. AT(c, r) = pSrc[c + r*8]. */
static void pq_rs_calc(Matrix44* P, Matrix44* Q, Matrix44* R, Matrix44* S, const jpgd_block_t* pSrc)
{
    const int a1[4] = { UF(0.415735f), UF(0.791065f), UF(-0.352443f), UF(0.277785f) };
    const int a2[4] = { UF(0.022887f), UF(-0.097545f), UF(0.490393f), UF(0.865723f) };
    const int b1[4] = { UF(0.906127f), UF(-0.318190f), UF(0.212608f), UF(-0.180240f) };
    const int b2[4] = { UF(-0.074658f), UF(0.513280f), UF(0.768178f), UF(-0.375330f) };
    int X0[4][8], X1[4][8];
#define AT(c, r) ((int)pSrc[(c) + (r) * 8])
    for (int j = 0; j < 8; ++j) {
        X0[0][j] = AT(0, j);
        X0[1][j] = UD(a1[0] * AT(1, j) + a1[1] * AT(3, j) + a1[2] * AT(5, j) + a1[3] * AT(7, j));
        X0[2][j] = AT(4, j);
        X0[3][j] = UD(a2[0] * AT(1, j) + a2[1] * AT(3, j) + a2[2] * AT(5, j) + a2[3] * AT(7, j));
        X1[0][j] = UD(b1[0] * AT(1, j) + b1[1] * AT(3, j) + b1[2] * AT(5, j) + b1[3] * AT(7, j));
        X1[1][j] = AT(2, j);
        X1[2][j] = UD(b2[0] * AT(1, j) + b2[1] * AT(3, j) + b2[2] * AT(5, j) + b2[3] * AT(7, j));
        X1[3][j] = AT(6, j);
    }
#undef AT
    for (int i = 0; i < 4; ++i) {
        const int* x = X0[i];
        P->v[i][0] = x[0];
        P->v[i][1] = UD(x[1] * a1[0] + x[3] * a1[1] + x[5] * a1[2] + x[7] * a1[3]);
        P->v[i][2] = x[4];
        P->v[i][3] = UD(x[1] * a2[0] + x[3] * a2[1] + x[5] * a2[2] + x[7] * a2[3]);
        Q->v[i][0] = UD(x[1] * b1[0] + x[3] * b1[1] + x[5] * b1[2] + x[7] * b1[3]);
        Q->v[i][1] = x[2];
        Q->v[i][2] = UD(x[1] * b2[0] + x[3] * b2[1] + x[5] * b2[2] + x[7] * b2[3]);
        Q->v[i][3] = x[6];
        x = X1[i];
        R->v[i][0] = x[0];
        R->v[i][1] = UD(x[1] * a1[0] + x[3] * a1[1] + x[5] * a1[2] + x[7] * a1[3]);
        R->v[i][2] = x[4];
        R->v[i][3] = UD(x[1] * a2[0] + x[3] * a2[1] + x[5] * a2[2] + x[7] * a2[3]);
        S->v[i][0] = UD(x[1] * b1[0] + x[3] * b1[1] + x[5] * b1[2] + x[7] * b1[3]);
        S->v[i][1] = x[2];
        S->v[i][2] = UD(x[1] * b2[0] + x[3] * b2[1] + x[5] * b2[2] + x[7] * b2[3]);
        S->v[i][3] = x[6];
    }
}
/* add_and_store / sub_and_store (jpegload.d:886-902): transposed store as short */
static void addsub_store(jpgd_block_t* pDst, const Matrix44* a, const Matrix44* b, int sub)
{
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c)
            pDst[c * 8 + r] = (jpgd_block_t)(sub ? a->v[r][c] - b->v[r][c] : a->v[r][c] + b->v[r][c]);
}

/* ---- decoder state ---- */
typedef struct {
    int valid;
    uint8_t num[17]; uint8_t val[256];
    /* canonical decode (Annex F.2.2.3): code -> symbol */
    int mincode[18], maxcode[18], valptr[18];
} hufftab;

typedef struct {
    const uint8_t* data; size_t len; size_t pos; int tem_flag;
    /* entropy bit reader */
    uint32_t bit_buf; int bits_left;
    int image_x_size, image_y_size, progressive;
    hufftab huff[MAX_HUFF_TABLES];
    int quant_valid[MAX_QUANT_TABLES]; jpgd_quant_t quant[MAX_QUANT_TABLES][64];
    int scan_type, comps_in_frame;
    int comp_h_samp[MAX_COMPONENTS], comp_v_samp[MAX_COMPONENTS], comp_quant[MAX_COMPONENTS], comp_ident[MAX_COMPONENTS];
    int comp_h_blocks[MAX_COMPONENTS], comp_v_blocks[MAX_COMPONENTS];
    int comps_in_scan, comp_list[MAX_COMPS_IN_SCAN], comp_dc_tab[MAX_COMPONENTS], comp_ac_tab[MAX_COMPONENTS];
    int max_mcu_x_size, max_mcu_y_size, blocks_per_mcu, max_blocks_per_row, mcus_per_row, mcus_per_col;
    int mcu_org[MAX_BLOCKS_PER_MCU];
    int total_lines_left, mcu_lines_left, real_dest_bytes_per_scan_line, dest_bytes_per_scan_line, dest_bytes_per_pixel;
    int restart_interval, restarts_left, next_restart_num;
    int max_mcus_per_row, max_blocks_per_mcu, expanded_blocks_per_mcu, expanded_blocks_per_row, expanded_blocks_per_component;
    int freq_domain_chroma_upsample, max_mcus_per_col;
    uint32_t last_dc_val[MAX_COMPONENTS];
    jpgd_block_t* pMCU_coefficients; int mcu_block_max_zag[MAX_BLOCKS_PER_MCU];
    uint8_t* pSample_buf;
    int crr[256], cbb[256], crg[256], cbg[256];
    uint8_t *pScan_line_0, *pScan_line_1;
    int error;
    float ppiX, ppiY, par;
    /* progressive (jpegload.d:440,470-476,3299-3683) */
    int spectral_start, spectral_end, successive_low, successive_high, eob_run;
    jpgd_block_t* dc_coeffs[MAX_COMPONENTS]; jpgd_block_t* ac_coeffs[MAX_COMPONENTS];
    int coef_num_x[MAX_COMPONENTS], coef_num_y[MAX_COMPONENTS];
    int block_y_mcu[MAX_COMPONENTS];
} jd;

/* get_char (jpegload.d:640-655): past the end, FF D9 FF D9 ... */
static unsigned get_char_pad(jd* d, int* pad)
{
    if (d->pos >= d->len) { if (pad) *pad = 1; int t = d->tem_flag; d->tem_flag ^= 1; return t ? 0xD9 : 0xFF; }
    if (pad) *pad = 0;
    return d->data[d->pos++];
}
static void stuff_char(jd* d) { d->pos--; }    /* only ever un-reads the byte just read */

/* get_octet (jpegload.d:683-696) */
static uint8_t get_octet(jd* d)
{
    int pad;
    unsigned c = get_char_pad(d, &pad);
    if (c == 0xFF) {
        if (pad) return 0xFF;
        size_t save = d->pos;
        c = get_char_pad(d, &pad);
        if (pad) { d->pos = save - 1; return 0xFF; }     /* stuff_char(0xFF) */
        if (c == 0x00) return 0xFF;
        d->pos = save - 1;                               /* stuff both back: the marker is never consumed */
        return 0xFF;
    }
    return (uint8_t)c;
}
/* get_bits (jpegload.d:699-719): markers are read as data */
static unsigned get_bits(jd* d, int num_bits)
{
    if (!num_bits) return 0;
    unsigned i = d->bit_buf >> (32 - num_bits);
    if ((d->bits_left -= num_bits) <= 0) {
        d->bit_buf <<= (num_bits += d->bits_left);
        unsigned c1 = get_char_pad(d, NULL);
        unsigned c2 = get_char_pad(d, NULL);
        d->bit_buf = (d->bit_buf & 0xFFFF0000u) | (c1 << 8) | c2;
        d->bit_buf <<= -d->bits_left;
        d->bits_left += 16;
    } else d->bit_buf <<= num_bits;
    return i;
}
/* get_bits_no_markers (jpegload.d:722-743) */
static unsigned get_bits_nm(jd* d, int num_bits)
{
    if (!num_bits) return 0;
    unsigned i = d->bit_buf >> (32 - num_bits);
    if ((d->bits_left -= num_bits) <= 0) {
        d->bit_buf <<= (num_bits += d->bits_left);
        unsigned c1 = get_octet(d);
        unsigned c2 = get_octet(d);
        d->bit_buf |= (c1 << 8) | c2;
        d->bit_buf <<= -d->bits_left;
        d->bits_left += 16;
    } else d->bit_buf <<= num_bits;
    return i;
}

/* make_huff_table (jpegload.d:2851-2987) restated as the canonical code it encodes */
static void make_huff_table(hufftab* h)
{
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
        h->valptr[l] = k;
        h->mincode[l] = code;
        code += h->num[l]; k += h->num[l];
        h->maxcode[l] = h->num[l] ? code - 1 : -1;
        code <<= 1;
    }
}
/* huff_decode (jpegload.d:746-813): returns the symbol, or -1 when the bits are not a code word */
static int huff_decode(jd* d, hufftab* h)
{
    unsigned top16 = d->bit_buf >> 16;
    for (int l = 1; l <= 16; ++l) {
        int code = (int)(top16 >> (16 - l));
        if (h->maxcode[l] >= 0 && code <= h->maxcode[l] && code >= h->mincode[l]) {
            int sym = h->val[h->valptr[l] + code - h->mincode[l]];
            get_bits_nm(d, l);
            return sym;
        }
    }
    return -1;
}
/* JPGD_HUFF_EXTEND (jpegload.d:816-822) */
static inline int huff_extend(int x, int s) { return (s == 0) ? x : ((x < (1 << (s - 1))) ? x + (int)(((unsigned)-1) << s) + 1 : x); }

/* ---- markers (jpegload.d:1177-1543) ---- */
static int read_dht_marker(jd* d)
{
    unsigned num_left = get_bits(d, 16);
    if (num_left < 2) return 0;
    num_left -= 2;
    while (num_left) {
        uint8_t huff_num[17], huff_val[256];
        int index = (int)get_bits(d, 8);
        huff_num[0] = 0;
        int count = 0;
        for (int i = 1; i <= 16; ++i) { huff_num[i] = (uint8_t)get_bits(d, 8); count += huff_num[i]; }
        if (count > 255) return 0;
        memset(huff_val, 0, sizeof(huff_val));
        for (int i = 0; i < count; ++i) huff_val[i] = (uint8_t)get_bits(d, 8);
        int i = 1 + 16 + count;
        if (num_left < (unsigned)i) return 0;
        num_left -= i;
        index = (index & 0x0F) + ((index & 0x10) >> 4) * (MAX_HUFF_TABLES >> 1);
        if (index >= MAX_HUFF_TABLES) return 0;
        d->huff[index].valid = 1;
        memcpy(d->huff[index].num, huff_num, 17);
        memcpy(d->huff[index].val, huff_val, 256);
    }
    return 1;
}
static int read_dqt_marker(jd* d)
{
    unsigned num_left = get_bits(d, 16);
    if (num_left < 2) return 0;
    num_left -= 2;
    while (num_left) {
        int n = (int)get_bits(d, 8);
        int prec = n >> 4;
        n &= 0x0F;
        if (n >= MAX_QUANT_TABLES) return 0;
        d->quant_valid[n] = 1;
        for (int i = 0; i < 64; ++i) {
            unsigned temp = get_bits(d, 8);
            if (prec) temp = (temp << 8) + get_bits(d, 8);
            d->quant[n][i] = (jpgd_quant_t)temp;
        }
        int i = 64 + 1;
        if (prec) i += 64;
        if (num_left < (unsigned)i) return 0;
        num_left -= i;
    }
    return 1;
}
static int read_sof_marker(jd* d)
{
    unsigned num_left = get_bits(d, 16);
    if (get_bits(d, 8) != 8) return 0;
    d->image_y_size = (int)get_bits(d, 16);
    if (d->image_y_size < 1 || d->image_y_size > MAX_HEIGHT) return 0;
    d->image_x_size = (int)get_bits(d, 16);
    if (d->image_x_size < 1 || d->image_x_size > MAX_WIDTH) return 0;
    d->comps_in_frame = (int)get_bits(d, 8);
    if (d->comps_in_frame > MAX_COMPONENTS) return 0;
    if (num_left != (unsigned)(d->comps_in_frame * 3 + 8)) return 0;
    for (int i = 0; i < d->comps_in_frame; ++i) {
        d->comp_ident[i] = (int)get_bits(d, 8);
        d->comp_h_samp[i] = (int)get_bits(d, 4);
        d->comp_v_samp[i] = (int)get_bits(d, 4);
        d->comp_quant[i] = (int)get_bits(d, 8);
    }
    return 1;
}
static int skip_variable_marker(jd* d)
{
    unsigned num_left = get_bits(d, 16);
    if (num_left < 2) return 0;
    num_left -= 2;
    while (num_left) { get_bits(d, 8); num_left--; }
    return 1;
}
static int read_dri_marker(jd* d)
{
    if (get_bits(d, 16) != 4) return 0;
    d->restart_interval = (int)get_bits(d, 16);
    return 1;
}
static int read_sos_marker(jd* d)
{
    unsigned num_left = get_bits(d, 16);
    int n = (int)get_bits(d, 8);
    d->comps_in_scan = n;
    num_left -= 3;
    if ((num_left != (unsigned)(n * 2 + 3)) || (n < 1) || (n > MAX_COMPS_IN_SCAN)) return 0;
    for (int i = 0; i < n; ++i) {
        int cc = (int)get_bits(d, 8);
        int c = (int)get_bits(d, 8);
        num_left -= 2;
        int ci;
        for (ci = 0; ci < d->comps_in_frame; ++ci) if (cc == d->comp_ident[ci]) break;
        if (ci >= d->comps_in_frame) return 0;
        d->comp_list[i] = ci;
        d->comp_dc_tab[ci] = (c >> 4) & 15;
        d->comp_ac_tab[ci] = (c & 15) + (MAX_HUFF_TABLES >> 1);
    }
    d->spectral_start = (int)get_bits(d, 8);
    d->spectral_end = (int)get_bits(d, 8);
    d->successive_high = (int)get_bits(d, 4);
    d->successive_low = (int)get_bits(d, 4);
    if (!d->progressive) { d->spectral_start = 0; d->spectral_end = 63; }       /* :1526-1530 */
    num_left -= 3;
    while (num_left) { get_bits(d, 8); num_left--; }
    return 1;
}
static int next_marker(jd* d)
{
    unsigned c;
    do {
        do { c = get_bits(d, 8); } while (c != 0xFF);
        do { c = get_bits(d, 8); } while (c == 0xFF);
    } while (c == 0);
    return (int)c;
}
static float inches_to_meters_x(float x) { return x / 39.37007874f; }   /* convertInchesToMeters, types.d:127 */

static uint16_t rd16(const uint8_t* p, int le) { return le ? (uint16_t)(p[0] | (p[1] << 8)) : (uint16_t)((p[0] << 8) | p[1]); }
static uint32_t rd32(const uint8_t* p, int le) { return le ? ((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24))
                                                           : (((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]); }

/* process_markers (jpegload.d:1578-1845). Returns the marker, or -1 on error. */
static int process_markers(jd* d, int allow_restarts)
{
    for (;;) {
        int c = next_marker(d);
        switch (c) {
        case M_SOF0: case M_SOF1: case M_SOF2: case M_SOF3: case M_SOF5: case M_SOF6: case M_SOF7:
        case M_SOF9: case M_SOF10: case M_SOF11: case M_SOF13: case M_SOF14: case M_SOF15:
        case M_SOI: case M_EOI: case M_SOS:
            return c;
        case M_DHT: if (!read_dht_marker(d)) return -1; break;
        case M_DAC: return -1;
        case M_DQT: if (!read_dqt_marker(d)) return -1; break;
        case M_DRI: if (!read_dri_marker(d)) return -1; break;
        case M_APP0: {
            unsigned num_left = get_bits(d, 16);
            if (num_left < 7) return -1;           /* the reference sets an error code here */
            num_left -= 2;
            uint8_t id[5];
            for (int i = 0; i < 5; ++i) id[i] = (uint8_t)get_bits(d, 8);
            num_left -= 5;
            static const uint8_t JFIF[5] = { 0x4A, 0x46, 0x49, 0x46, 0x00 };
            if (memcmp(id, JFIF, 5) == 0 && num_left >= 7) {
                get_bits(d, 16);
                unsigned units = get_bits(d, 8);
                int Xd = (int)get_bits(d, 16), Yd = (int)get_bits(d, 16);
                num_left -= 7;
                d->par = (float)(Xd / (double)Yd);
                switch (units) {
                case 0: d->ppiX = -1; d->ppiY = -1; break;
                case 1: d->ppiX = (float)Xd; d->ppiY = (float)Yd; break;
                case 2: d->ppiX = inches_to_meters_x(Xd * 100.0f); d->ppiY = inches_to_meters_x(Yd * 100.0f); break;
                default: break;
                }
            }
            while (num_left) { get_bits(d, 8); num_left--; }
            break; }
        case M_APP0 + 1: {
            unsigned num_left = get_bits(d, 16);
            if (num_left < 2) return -1;
            num_left -= 2;
            uint8_t* ex = (uint8_t*)malloc(num_left ? num_left : 1);
            for (unsigned i = 0; i < num_left; ++i) ex[i] = (uint8_t)get_bits(d, 8);
            static const uint8_t EXIF[6] = { 0x45, 0x78, 0x69, 0x66, 0x00, 0x00 };
            int bad = 0;
            if (num_left >= 14 && memcmp(ex, EXIF, 6) == 0) {
                const uint8_t* tiff = ex + 6; unsigned tlen = num_left - 6;
                uint16_t bo = rd16(tiff, 0);
                if (bo != 0x4949 && bo != 0x4D4D) bad = 1;
                else {
                    int le = bo == 0x4949;
                    if (rd16(tiff + 2, le) != 42) bad = 1;
                    else {
                        uint32_t offset = rd32(tiff + 4, le);
                        double rx = 72, ry = 72; int unit = 2;
                        while (offset != 0 && !bad) {
                            if (offset > num_left || offset + 2 > tlen) { bad = 1; break; }
                            const uint8_t* p = tiff + offset;
                            unsigned ne = rd16(p, le); p += 2;
                            if ((size_t)(p - tiff) + (size_t)ne * 12 + 4 > tlen) { bad = 1; break; }
                            for (unsigned e = 0; e < ne; ++e, p += 12) {
                                unsigned tag = rd16(p, le); uint32_t vo = rd32(p + 8, le);
                                if (tag == 282 || tag == 283) {
                                    if ((size_t)vo + 8 > tlen) { bad = 1; break; }
                                    double num = rd32(tiff + vo, le), den = rd32(tiff + vo + 4, le);
                                    if (tag == 282) rx = num / den; else ry = num / den;
                                }
                                if (tag == 296) unit = (int)vo;
                            }
                            if (bad) break;
                            offset = rd32(p, le);
                        }
                        if (!bad) {
                            if (unit == 2) { d->ppiX = (float)rx; d->ppiY = (float)ry; d->par = (float)(rx / ry); }
                            else if (unit == 3) { d->ppiX = inches_to_meters_x((float)(rx * 100)); d->ppiY = inches_to_meters_x((float)(ry * 100)); d->par = (float)(rx / ry); }
                        }
                    }
                }
            }
            free(ex);
            if (bad) return -1;
            break; }
        case 0xD0: case 0xD1: case 0xD2: case 0xD3: case 0xD4: case 0xD5: case 0xD6: case 0xD7:
            if (allow_restarts) continue;
            return -1;
        case M_JPG: case M_TEM:
            return -1;
        default:
            if (!skip_variable_marker(d)) return -1;
            break;
        }
    }
}

/* locate_soi_marker (jpegload.d:1851-1895) */
static int locate_soi_marker(jd* d)
{
    unsigned lastchar = get_bits(d, 8), thischar = get_bits(d, 8);
    if (lastchar == 0xFF && thischar == M_SOI) return 1;
    unsigned bytesleft = 4096;
    for (;;) {
        if (--bytesleft == 0) return 0;
        lastchar = thischar;
        thischar = get_bits(d, 8);
        if (lastchar == 0xFF) {
            if (thischar == M_SOI) break;
            else if (thischar == M_EOI) return 0;
        }
    }
    thischar = (d->bit_buf >> 24) & 0xFF;
    if (thischar != 0xFF) return 0;
    return 1;
}

/* calc_mcu_block_order (jpegload.d:3038-3090) */
static void calc_mcu_block_order(jd* d)
{
    int max_h = 0, max_v = 0;
    for (int c = 0; c < d->comps_in_frame; ++c) {
        if (d->comp_h_samp[c] > max_h) max_h = d->comp_h_samp[c];
        if (d->comp_v_samp[c] > max_v) max_v = d->comp_v_samp[c];
    }
    for (int c = 0; c < d->comps_in_frame; ++c) {
        d->comp_h_blocks[c] = ((((d->image_x_size * d->comp_h_samp[c]) + (max_h - 1)) / max_h) + 7) / 8;
        d->comp_v_blocks[c] = ((((d->image_y_size * d->comp_v_samp[c]) + (max_v - 1)) / max_v) + 7) / 8;
    }
    if (d->comps_in_scan == 1) {
        d->mcus_per_row = d->comp_h_blocks[d->comp_list[0]];
        d->mcus_per_col = d->comp_v_blocks[d->comp_list[0]];
        d->mcu_org[0] = d->comp_list[0];
        d->blocks_per_mcu = 1;
    } else {
        d->mcus_per_row = (((d->image_x_size + 7) / 8) + (max_h - 1)) / max_h;
        d->mcus_per_col = (((d->image_y_size + 7) / 8) + (max_v - 1)) / max_v;
        d->blocks_per_mcu = 0;
        for (int n = 0; n < d->comps_in_scan; ++n) {
            int c = d->comp_list[n];
            int nb = d->comp_h_samp[c] * d->comp_v_samp[c];
            while (nb--) { if (d->blocks_per_mcu >= MAX_BLOCKS_PER_MCU) { d->error = 1; return; } d->mcu_org[d->blocks_per_mcu++] = c; }
        }
    }
}

/* create_look_ups (jpegload.d:2080-2094); FIX!(x) = (int)(x * 65536 + 0.5f) */
static void create_look_ups(jd* d)
{
    const int F140200 = (int)(1.40200f * 65536 + 0.5f), F177200 = (int)(1.77200f * 65536 + 0.5f);
    const int F071414 = (int)(0.71414f * 65536 + 0.5f), F034414 = (int)(0.34414f * 65536 + 0.5f);
    for (int i = 0; i <= 255; ++i) {
        int k = i - 128;
        d->crr[i] = (F140200 * k + 32768) >> 16;
        d->cbb[i] = (F177200 * k + 32768) >> 16;
        d->crg[i] = (-F071414) * k;
        d->cbg[i] = (-F034414) * k + 32768;
    }
}
/* the pixel expression of H1V1Convert / H2V1Convert / H1V2Convert / expanded_convert (jpegload.d:2536-2549 ...):
 * saturating packs == clamp for these ranges */
static inline void ycc(jd* d, int y, int cb, int cr, uint8_t* o)
{
    o[0] = CLAMP(y + d->crr[cr]);
    o[1] = CLAMP(y + ((d->crg[cr] + d->cbg[cb]) >> 16));
    o[2] = CLAMP(y + d->cbb[cb]);
    o[3] = 255;
}

/* init_frame (jpegload.d:3130-3268) */
static int init_frame(jd* d)
{
    if (d->comps_in_frame == 1) {
        if (d->comp_h_samp[0] != 1 || d->comp_v_samp[0] != 1) return 0;
        d->scan_type = GRAYSCALE; d->max_blocks_per_mcu = 1; d->max_mcu_x_size = 8; d->max_mcu_y_size = 8;
    } else if (d->comps_in_frame == 3) {
        if (d->comp_h_samp[1] != 1 || d->comp_v_samp[1] != 1 || d->comp_h_samp[2] != 1 || d->comp_v_samp[2] != 1) return 0;
        if (d->comp_h_samp[0] == 1 && d->comp_v_samp[0] == 1) { d->scan_type = YH1V1; d->max_blocks_per_mcu = 3; d->max_mcu_x_size = 8; d->max_mcu_y_size = 8; }
        else if (d->comp_h_samp[0] == 2 && d->comp_v_samp[0] == 1) { d->scan_type = YH2V1; d->max_blocks_per_mcu = 4; d->max_mcu_x_size = 16; d->max_mcu_y_size = 8; }
        else if (d->comp_h_samp[0] == 1 && d->comp_v_samp[0] == 2) { d->scan_type = YH1V2; d->max_blocks_per_mcu = 4; d->max_mcu_x_size = 8; d->max_mcu_y_size = 16; }
        else if (d->comp_h_samp[0] == 2 && d->comp_v_samp[0] == 2) { d->scan_type = YH2V2; d->max_blocks_per_mcu = 6; d->max_mcu_x_size = 16; d->max_mcu_y_size = 16; }
        else return 0;
    } else return 0;
    d->max_mcus_per_row = (d->image_x_size + (d->max_mcu_x_size - 1)) / d->max_mcu_x_size;
    d->max_mcus_per_col = (d->image_y_size + (d->max_mcu_y_size - 1)) / d->max_mcu_y_size;
    d->dest_bytes_per_pixel = d->scan_type == GRAYSCALE ? 1 : 4;
    d->dest_bytes_per_scan_line = ((d->image_x_size + 15) & 0xFFF0) * d->dest_bytes_per_pixel;
    d->real_dest_bytes_per_scan_line = d->image_x_size * d->dest_bytes_per_pixel;
    d->pScan_line_0 = (uint8_t*)calloc((size_t)d->dest_bytes_per_scan_line + 64, 1);
    d->pScan_line_1 = (uint8_t*)calloc((size_t)d->dest_bytes_per_scan_line + 64, 1);
    d->max_blocks_per_row = d->max_mcus_per_row * d->max_blocks_per_mcu;
    if (d->max_blocks_per_row > MAX_BLOCKS_PER_ROW) return 0;
    d->pMCU_coefficients = (jpgd_block_t*)calloc((size_t)d->max_blocks_per_mcu * 64, sizeof(jpgd_block_t));
    for (int i = 0; i < d->max_blocks_per_mcu; ++i) d->mcu_block_max_zag[i] = 64;
    d->expanded_blocks_per_component = d->comp_h_samp[0] * d->comp_v_samp[0];
    d->expanded_blocks_per_mcu = d->expanded_blocks_per_component * d->comps_in_frame;
    d->expanded_blocks_per_row = d->max_mcus_per_row * d->expanded_blocks_per_mcu;
    d->freq_domain_chroma_upsample = (d->expanded_blocks_per_mcu == 4 * 3);   /* version JPGD_SUPPORT_FREQ_DOMAIN_UPSAMPLING, :59 */
    size_t nb = (size_t)(d->freq_domain_chroma_upsample ? d->expanded_blocks_per_row : d->max_blocks_per_row) * 64;
    d->pSample_buf = (uint8_t*)calloc(nb + 64, 1);
    d->total_lines_left = d->image_y_size;
    d->mcu_lines_left = 0;
    create_look_ups(d);
    return 1;
}

/* ---- test hooks (tests/test_oracle_reference_text.py): the arithmetic kernels of this file, callable on
 * single blocks so that they can be compared with vectors generated from the reference's source text
 * (tests/golden/gen_from_reference.py). ---- */
void or_test_jpeg_idct(const int16_t* src, int max_zag, uint8_t* dst) { idct(src, dst, max_zag); }
void or_test_jpeg_upsample(const int16_t* src, uint8_t* dst256)          /* chroma half of transform_mcu_expand */
{
    Matrix44 P, Q, R, S, a, bb, c, dd;
    jpgd_block_t temp_block[64];
    memset(temp_block, 0, sizeof(temp_block));
    pq_rs_calc(&P, &Q, &R, &S, src);
    for (int r = 0; r < 4; ++r) for (int q = 0; q < 4; ++q) {
        a.v[r][q] = P.v[r][q] + Q.v[r][q]; bb.v[r][q] = P.v[r][q] - Q.v[r][q];
        c.v[r][q] = R.v[r][q] + S.v[r][q]; dd.v[r][q] = R.v[r][q] - S.v[r][q];
    }
    addsub_store(temp_block, &a, &c, 0);  idct_4x4(temp_block, dst256);
    addsub_store(temp_block, &a, &c, 1);  idct_4x4(temp_block, dst256 + 64);
    addsub_store(temp_block, &bb, &dd, 0); idct_4x4(temp_block, dst256 + 128);
    addsub_store(temp_block, &bb, &dd, 1); idct_4x4(temp_block, dst256 + 192);
}
void or_test_jpeg_ycc(int y, int cb, int cr, uint8_t* rgb, int* tables /* crr, cbb, crg, cbg: 4 x 256, may be NULL */)
{
    jd* d = (jd*)calloc(1, sizeof(jd));
    uint8_t o[4];
    create_look_ups(d);
    ycc(d, y, cb, cr, o);
    rgb[0] = o[0]; rgb[1] = o[1]; rgb[2] = o[2];
    if (tables) { memcpy(tables, d->crr, 1024); memcpy(tables + 256, d->cbb, 1024); memcpy(tables + 512, d->crg, 1024); memcpy(tables + 768, d->cbg, 1024); }
    free(d);
}

/* fix_in_buffer + init_scan (jpegload.d:2098-2118, 3093-3127) */
/* init_scan (jpegload.d:3090-3128): 1 = a scan is ready, 0 = EOI (no further scan), -1 = error */
static int init_scan3(jd* d)
{
    int c = process_markers(d, 0);
    if (c < 0) return -1;
    if (c == M_EOI) return 0;
    if (c != M_SOS) return -1;
    if (!read_sos_marker(d)) return -1;
    calc_mcu_block_order(d);
    if (d->error) return -1;
    /* check_huff_tables / check_quant_tables (:2990-3035): a DC table is needed when the scan covers coefficient 0,
     * an AC table when it covers any other (the reference records the error and fails later; here: at once) */
    for (int i = 0; i < d->comps_in_scan; ++i) {
        int ci = d->comp_list[i];
        if (d->spectral_start == 0 && (d->comp_dc_tab[ci] >= MAX_HUFF_TABLES || !d->huff[d->comp_dc_tab[ci]].valid)) return -1;
        if (d->spectral_end > 0 && (d->comp_ac_tab[ci] >= MAX_HUFF_TABLES || !d->huff[d->comp_ac_tab[ci]].valid)) return -1;
        if (d->comp_quant[ci] >= MAX_QUANT_TABLES || !d->quant_valid[d->comp_quant[ci]]) return -1;
    }
    for (int i = 0; i < MAX_HUFF_TABLES; ++i) if (d->huff[i].valid) make_huff_table(&d->huff[i]);
    memset(d->last_dc_val, 0, sizeof(d->last_dc_val));
    d->eob_run = 0;
    if (d->restart_interval) { d->restarts_left = d->restart_interval; d->next_restart_num = 0; }
    /* fix_in_buffer: un-read what the marker scanner pre-fetched, then prime with marker-aware reads */
    {
        size_t n = 2 + (d->bits_left >= 8) + (d->bits_left == 16);
        if (d->pos < n || d->pos > d->len) return -1;     /* SOS at the very end of the data */
        d->pos -= n;
    }
    d->bits_left = 16;
    get_bits_nm(d, 16);
    get_bits_nm(d, 16);
    return 1;
}
static int init_scan(jd* d) { return init_scan3(d) == 1; }

/* process_restart (jpegload.d:2335-2402) */
static int process_restart(jd* d)
{
    int i, c = 0;
    for (i = 1536; i > 0; i--) if (get_char_pad(d, NULL) == 0xFF) break;
    if (i == 0) return 0;
    for (; i > 0; i--) if ((c = (int)get_char_pad(d, NULL)) != 0xFF) break;
    if (i == 0) return 0;
    if (c != (d->next_restart_num + M_RST0)) return 0;
    memset(d->last_dc_val, 0, (size_t)d->comps_in_frame * sizeof(uint32_t));
    d->eob_run = 0;
    d->restarts_left = d->restart_interval;
    d->next_restart_num = (d->next_restart_num + 1) & 7;
    d->bits_left = 16;
    get_bits_nm(d, 16);
    get_bits_nm(d, 16);
    return 1;
}

/* transform_mcu / transform_mcu_expand (jpegload.d:2120-2255) */
static void transform_mcu(jd* d, int mcu_row)
{
    jpgd_block_t* pSrc = d->pMCU_coefficients;
    uint8_t* pDst = d->pSample_buf + (size_t)mcu_row * d->blocks_per_mcu * 64;
    for (int b = 0; b < d->blocks_per_mcu; ++b) { idct(pSrc, pDst, d->mcu_block_max_zag[b]); pSrc += 64; pDst += 64; }
}
static void transform_mcu_expand(jd* d, int mcu_row)
{
    jpgd_block_t* pSrc = d->pMCU_coefficients;
    uint8_t* pDst = d->pSample_buf + (size_t)mcu_row * d->expanded_blocks_per_mcu * 64;
    int b;
    for (b = 0; b < d->expanded_blocks_per_component; ++b) { idct(pSrc, pDst, d->mcu_block_max_zag[b]); pSrc += 64; pDst += 64; }
    jpgd_block_t temp_block[64];
    for (int i = 0; i < 2; ++i) {
        Matrix44 P, Q, R, S, a, bb, c, dd;
        memset(temp_block, 0, sizeof(temp_block));
        pq_rs_calc(&P, &Q, &R, &S, pSrc);
        for (int r = 0; r < 4; ++r) for (int q = 0; q < 4; ++q) {
            a.v[r][q] = P.v[r][q] + Q.v[r][q]; bb.v[r][q] = P.v[r][q] - Q.v[r][q];
            c.v[r][q] = R.v[r][q] + S.v[r][q]; dd.v[r][q] = R.v[r][q] - S.v[r][q];
        }
        addsub_store(temp_block, &a, &c, 0);  idct_4x4(temp_block, pDst); pDst += 64;
        addsub_store(temp_block, &a, &c, 1);  idct_4x4(temp_block, pDst); pDst += 64;
        addsub_store(temp_block, &bb, &dd, 0); idct_4x4(temp_block, pDst); pDst += 64;
        addsub_store(temp_block, &bb, &dd, 1); idct_4x4(temp_block, pDst); pDst += 64;
        pSrc += 64;
    }
}

/* decode_next_row (jpegload.d:2405-2525) */
static int decode_next_row(jd* d)
{
    for (int mcu_row = 0; mcu_row < d->mcus_per_row; ++mcu_row) {
        if (d->restart_interval && d->restarts_left == 0) { if (!process_restart(d)) return 0; }
        jpgd_block_t* p = d->pMCU_coefficients;
        for (int mcu_block = 0; mcu_block < d->blocks_per_mcu; ++mcu_block, p += 64) {
            int component_id = d->mcu_org[mcu_block];
            jpgd_quant_t* q = d->quant[d->comp_quant[component_id]];
            int r, s;
            s = huff_decode(d, &d->huff[d->comp_dc_tab[component_id]]);
            if (s < 0 || s > 15) return 0;
            r = (int)get_bits_nm(d, s & 0xF);
            s = huff_extend(r, s);
            d->last_dc_val[component_id] = (uint32_t)(s += (int)d->last_dc_val[component_id]);
            p[0] = (jpgd_block_t)(s * q[0]);
            int prev_num_set = d->mcu_block_max_zag[mcu_block];
            hufftab* pH = &d->huff[d->comp_ac_tab[component_id]];
            int k;
            for (k = 1; k < 64; ++k) {
                s = huff_decode(d, pH);
                if (s < 0) return 0;
                int extra_bits = (int)get_bits_nm(d, s & 0xF);
                r = s >> 4;
                s &= 15;
                if (s) {
                    if (r) {
                        if ((k + r) > 63) return 0;
                        if (k < prev_num_set) { int n = r < prev_num_set - k ? r : prev_num_set - k; int kt = k; while (n--) p[g_ZAG[kt++]] = 0; }
                        k += r;
                    }
                    s = huff_extend(extra_bits, s);
                    p[g_ZAG[k]] = (jpgd_block_t)(s * q[k]);
                } else {
                    if (r == 15) {
                        if ((k + 16) > 64) return 0;
                        if (k < prev_num_set) { int n = 16 < prev_num_set - k ? 16 : prev_num_set - k; int kt = k; while (n--) p[g_ZAG[kt++]] = 0; }
                        k += 16 - 1;
                    } else break;
                }
            }
            if (k < prev_num_set) { int kt = k; while (kt < prev_num_set) p[g_ZAG[kt++]] = 0; }
            d->mcu_block_max_zag[mcu_block] = k;
        }
        if (d->freq_domain_chroma_upsample) transform_mcu_expand(d, mcu_row);
        else transform_mcu(d, mcu_row);
        d->restarts_left--;
    }
    return 1;
}

/* ---- progressive (SOF2): every scan decoded into whole-image coefficient buffers, then one MCU row at a time is
 * dequantised and transformed (jpegload.d:3274-3683, :2259-2332) ---- */
static jpgd_block_t* dc_getp(jd* d, int c, int bx, int by) { return d->dc_coeffs[c] + ((size_t)by * d->coef_num_x[c] + bx); }
static jpgd_block_t* ac_getp(jd* d, int c, int bx, int by) { return d->ac_coeffs[c] + ((size_t)by * d->coef_num_x[c] + bx) * 64; }
static inline int shl(int v, int n) { return (int)((unsigned)v << n); }

/* decode_block_dc_first (:3299-3320) */
static int decode_block_dc_first(jd* d, int c, int bx, int by)
{
    jpgd_block_t* p = dc_getp(d, c, bx, by);
    int s = huff_decode(d, &d->huff[d->comp_dc_tab[c]]);
    if (s < 0 || s > 15) return 0;
    if (s != 0) { int r = (int)get_bits_nm(d, s); s = huff_extend(r, s); }
    d->last_dc_val[c] = (uint32_t)(s += (int)d->last_dc_val[c]);
    p[0] = (jpgd_block_t)shl(s, d->successive_low);
    return 1;
}
/* decode_block_dc_refine (:3322-3333) */
static int decode_block_dc_refine(jd* d, int c, int bx, int by)
{
    if (get_bits_nm(d, 1)) { jpgd_block_t* p = dc_getp(d, c, bx, by); p[0] = (jpgd_block_t)(p[0] | (1 << d->successive_low)); }
    return 1;
}
/* decode_block_ac_first (:3335-3398) */
static int decode_block_ac_first(jd* d, int c, int bx, int by)
{
    if (d->eob_run) { d->eob_run--; return 1; }
    jpgd_block_t* p = ac_getp(d, c, bx, by);
    for (int k = d->spectral_start; k <= d->spectral_end; k++) {
        int s = huff_decode(d, &d->huff[d->comp_ac_tab[c]]);
        if (s < 0) return 0;
        int r = s >> 4;
        s &= 15;
        if (s) {
            if ((k += r) > 63) return 0;
            r = (int)get_bits_nm(d, s);
            s = huff_extend(r, s);
            p[g_ZAG[k]] = (jpgd_block_t)shl(s, d->successive_low);
        } else {
            if (r == 15) { if ((k += 15) > 63) return 0; }
            else {
                d->eob_run = 1 << r;
                if (r) d->eob_run += (int)get_bits_nm(d, r);
                d->eob_run--;
                break;
            }
        }
    }
    return 1;
}
/* decode_block_ac_refine (:3400-3519) */
static int decode_block_ac_refine(jd* d, int c, int bx, int by)
{
    const int p1 = 1 << d->successive_low, m1 = shl(-1, d->successive_low);
    jpgd_block_t* p = ac_getp(d, c, bx, by);
    int k = d->spectral_start;
    if (d->eob_run == 0) {
        for (; k <= d->spectral_end; k++) {
            int s = huff_decode(d, &d->huff[d->comp_ac_tab[c]]);
            if (s < 0) return 0;
            int r = s >> 4;
            s &= 15;
            if (s) {
                if (s != 1) return 0;
                s = get_bits_nm(d, 1) ? p1 : m1;
            } else if (r != 15) {
                d->eob_run = 1 << r;
                if (r) d->eob_run += (int)get_bits_nm(d, r);
                break;
            }
            do {
                jpgd_block_t* this_coef = p + g_ZAG[k & 63];
                if (*this_coef != 0) {
                    if (get_bits_nm(d, 1)) {
                        if ((*this_coef & p1) == 0) *this_coef = (jpgd_block_t)(*this_coef + (*this_coef >= 0 ? p1 : m1));
                    }
                } else if (--r < 0) break;
                k++;
            } while (k <= d->spectral_end);
            if (s && k < 64) p[g_ZAG[k]] = (jpgd_block_t)s;
        }
    }
    if (d->eob_run > 0) {
        for (; k <= d->spectral_end; k++) {
            jpgd_block_t* this_coef = p + g_ZAG[k & 63];
            if (*this_coef != 0) {
                if (get_bits_nm(d, 1)) {
                    if ((*this_coef & p1) == 0) *this_coef = (jpgd_block_t)(*this_coef + (*this_coef >= 0 ? p1 : m1));
                }
            }
        }
        d->eob_run--;
    }
    return 1;
}
typedef int (*decode_block_fn)(jd*, int, int, int);
/* decode_scan (:3521-3585) */
static int decode_scan(jd* d, decode_block_fn fn)
{
    int block_x_mcu[MAX_COMPONENTS], block_y_mcu[MAX_COMPONENTS];
    memset(block_y_mcu, 0, sizeof(block_y_mcu));
    for (int mcu_col = 0; mcu_col < d->mcus_per_col; mcu_col++) {
        memset(block_x_mcu, 0, sizeof(block_x_mcu));
        for (int mcu_row = 0; mcu_row < d->mcus_per_row; mcu_row++) {
            int xo = 0, yo = 0;
            if (d->restart_interval && d->restarts_left == 0) { if (!process_restart(d)) return 0; }
            for (int mcu_block = 0; mcu_block < d->blocks_per_mcu; mcu_block++) {
                const int c = d->mcu_org[mcu_block];
                if (block_x_mcu[c] + xo >= d->coef_num_x[c] || block_y_mcu[c] + yo >= d->coef_num_y[c]) return 0;   /* the reference asserts */
                if (!fn(d, c, block_x_mcu[c] + xo, block_y_mcu[c] + yo)) return 0;
                if (d->comps_in_scan == 1) block_x_mcu[c]++;
                else if (++xo == d->comp_h_samp[c]) {
                    xo = 0;
                    if (++yo == d->comp_v_samp[c]) { yo = 0; block_x_mcu[c] += d->comp_h_samp[c]; }
                }
            }
            d->restarts_left--;
        }
        if (d->comps_in_scan == 1) block_y_mcu[d->comp_list[0]]++;
        else for (int n = 0; n < d->comps_in_scan; n++) { const int c = d->comp_list[n]; block_y_mcu[c] += d->comp_v_samp[c]; }
    }
    return 1;
}
/* init_progressive (:3587-3684) */
static int init_progressive(jd* d)
{
    if (d->comps_in_frame == 4) return 0;
    for (int i = 0; i < d->comps_in_frame; i++) {
        d->coef_num_x[i] = d->max_mcus_per_row * d->comp_h_samp[i];
        d->coef_num_y[i] = d->max_mcus_per_col * d->comp_v_samp[i];
        const size_t nb = (size_t)d->coef_num_x[i] * d->coef_num_y[i];
        d->dc_coeffs[i] = (jpgd_block_t*)calloc(nb, sizeof(jpgd_block_t));
        d->ac_coeffs[i] = (jpgd_block_t*)calloc(nb * 64, sizeof(jpgd_block_t));
        if (!d->dc_coeffs[i] || !d->ac_coeffs[i]) return 0;
    }
    for (;;) {
        const int si = init_scan3(d);
        if (si < 0) return 0;
        if (!si) break;
        const int dc_only_scan = d->spectral_start == 0, refinement_scan = d->successive_high != 0;
        if (d->spectral_start > d->spectral_end || d->spectral_end > 63) return 0;
        if (dc_only_scan) { if (d->spectral_end) return 0; }
        else if (d->comps_in_scan != 1) return 0;
        if (refinement_scan && d->successive_low != d->successive_high - 1) return 0;
        decode_block_fn fn = dc_only_scan ? (refinement_scan ? decode_block_dc_refine : decode_block_dc_first)
                                          : (refinement_scan ? decode_block_ac_refine : decode_block_ac_first);
        if (!decode_scan(d, fn)) return 0;
        d->bits_left = 16;
        get_bits(d, 16);
        get_bits(d, 16);
    }
    d->comps_in_scan = d->comps_in_frame;
    for (int i = 0; i < d->comps_in_frame; i++) d->comp_list[i] = i;
    calc_mcu_block_order(d);
    return !d->error;
}
/* load_next_row (:2259-2332) */
static void load_next_row(jd* d)
{
    int block_x_mcu[MAX_COMPONENTS];
    memset(block_x_mcu, 0, sizeof(block_x_mcu));
    for (int mcu_row = 0; mcu_row < d->mcus_per_row; mcu_row++) {
        int xo = 0, yo = 0;
        for (int mcu_block = 0; mcu_block < d->blocks_per_mcu; mcu_block++) {
            const int c = d->mcu_org[mcu_block];
            const jpgd_quant_t* q = d->quant[d->comp_quant[c]];
            jpgd_block_t* p = d->pMCU_coefficients + 64 * mcu_block;
            const jpgd_block_t* pAC = ac_getp(d, c, block_x_mcu[c] + xo, d->block_y_mcu[c] + yo);
            const jpgd_block_t* pDC = dc_getp(d, c, block_x_mcu[c] + xo, d->block_y_mcu[c] + yo);
            p[0] = pDC[0];
            memcpy(&p[1], &pAC[1], 63 * sizeof(jpgd_block_t));
            int i;
            for (i = 63; i > 0; i--) if (p[g_ZAG[i]]) break;
            d->mcu_block_max_zag[mcu_block] = i + 1;
            for (; i >= 0; i--) if (p[g_ZAG[i]]) p[g_ZAG[i]] = (jpgd_block_t)(p[g_ZAG[i]] * q[i]);
            if (d->comps_in_scan == 1) block_x_mcu[c]++;
            else if (++xo == d->comp_h_samp[c]) {
                xo = 0;
                if (++yo == d->comp_v_samp[c]) { yo = 0; block_x_mcu[c] += d->comp_h_samp[c]; }
            }
        }
        if (d->freq_domain_chroma_upsample) transform_mcu_expand(d, mcu_row);
        else transform_mcu(d, mcu_row);
    }
    if (d->comps_in_scan == 1) d->block_y_mcu[d->comp_list[0]]++;
    else for (int n = 0; n < d->comps_in_scan; n++) { const int c = d->comp_list[n]; d->block_y_mcu[c] += d->comp_v_samp[c]; }
}

/* colour conversion (jpegload.d:2528-2823) */
static void H1V1Convert(jd* d)
{
    int row = d->max_mcu_y_size - d->mcu_lines_left;
    uint8_t* o = d->pScan_line_0; uint8_t* s = d->pSample_buf + row * 8;
    for (int i = d->max_mcus_per_row; i > 0; i--) {
        for (int j = 0; j < 8; ++j) { ycc(d, s[j], s[64 + j], s[128 + j], o); o += 4; }
        s += 64 * 3;
    }
}
static void H2V1Convert(jd* d)
{
    int row = d->max_mcu_y_size - d->mcu_lines_left;
    uint8_t* d0 = d->pScan_line_0; uint8_t* y = d->pSample_buf + row * 8; uint8_t* c = d->pSample_buf + 2 * 64 + row * 8;
    for (int i = d->max_mcus_per_row; i > 0; i--) {
        for (int l = 0; l < 2; ++l) {
            for (int j = 0; j < 4; ++j) {
                int cb = c[0], cr = c[64];
                ycc(d, y[j << 1], cb, cr, d0); ycc(d, y[(j << 1) + 1], cb, cr, d0 + 4);
                d0 += 8; c++;
            }
            y += 64;
        }
        y += 64 * 4 - 64 * 2;
        c += 64 * 4 - 8;
    }
}
static void H1V2Convert(jd* d)
{
    int row = d->max_mcu_y_size - d->mcu_lines_left;
    uint8_t *d0 = d->pScan_line_0, *d1 = d->pScan_line_1, *y, *c;
    if (row < 8) y = d->pSample_buf + row * 8; else y = d->pSample_buf + 64 * 1 + (row & 7) * 8;
    c = d->pSample_buf + 64 * 2 + (row >> 1) * 8;
    for (int i = d->max_mcus_per_row; i > 0; i--) {
        for (int j = 0; j < 8; ++j) {
            int cb = c[0 + j], cr = c[64 + j];
            ycc(d, y[j], cb, cr, d0); ycc(d, y[8 + j], cb, cr, d1);
            d0 += 4; d1 += 4;
        }
        y += 64 * 4; c += 64 * 4;
    }
}
static void gray_convert(jd* d)
{
    int row = d->max_mcu_y_size - d->mcu_lines_left;
    uint8_t* o = d->pScan_line_0; uint8_t* s = d->pSample_buf + row * 8;
    for (int i = d->max_mcus_per_row; i > 0; i--) { memcpy(o, s, 8); s += 64; o += 8; }
}
static void expanded_convert(jd* d)
{
    int row = d->max_mcu_y_size - d->mcu_lines_left;
    uint8_t* Py = d->pSample_buf + (row / 8) * 64 * d->comp_h_samp[0] + (row & 7) * 8;
    uint8_t* o = d->pScan_line_0;
    for (int i = d->max_mcus_per_row; i > 0; i--) {
        for (int k = 0; k < d->max_mcu_x_size; k += 8) {
            const int Y_ofs = k * 8;
            const int Cb_ofs = Y_ofs + 64 * d->expanded_blocks_per_component;
            const int Cr_ofs = Y_ofs + 64 * d->expanded_blocks_per_component * 2;
            for (int j = 0; j < 8; ++j) { ycc(d, Py[Y_ofs + j], Py[Cb_ofs + j], Py[Cr_ofs + j], o); o += 4; }
        }
        Py += 64 * d->expanded_blocks_per_mcu;
    }
}

/* decode (jpegload.d:545-612). H2V2 always takes the expanded path (freq-domain upsampling on). */
static int decode_line(jd* d, const uint8_t** pScan_line)
{
    if (d->total_lines_left == 0) return 1;
    if (d->mcu_lines_left == 0) {
        if (d->progressive) load_next_row(d);              /* decode (:555-563) */
        else if (!decode_next_row(d)) return -1;
        d->mcu_lines_left = d->max_mcu_y_size;
    }
    if (d->freq_domain_chroma_upsample) { expanded_convert(d); *pScan_line = d->pScan_line_0; }
    else switch (d->scan_type) {
        case YH2V1: H2V1Convert(d); *pScan_line = d->pScan_line_0; break;
        case YH1V2:
            if ((d->mcu_lines_left & 1) == 0) { H1V2Convert(d); *pScan_line = d->pScan_line_0; }
            else *pScan_line = d->pScan_line_1;
            break;
        case YH1V1: H1V1Convert(d); *pScan_line = d->pScan_line_0; break;
        case GRAYSCALE: gray_convert(d); *pScan_line = d->pScan_line_0; break;
        default: return -1;
    }
    --d->mcu_lines_left;
    --d->total_lines_left;
    return 0;
}

static void jd_free(jd* d)
{
    free(d->pScan_line_0); free(d->pScan_line_1); free(d->pMCU_coefficients); free(d->pSample_buf);
    for (int i = 0; i < MAX_COMPONENTS; ++i) { free(d->dc_coeffs[i]); free(d->ac_coeffs[i]); }
}

/* decompress_jpeg_image_from_stream (jpegload.d:3720-3808) */
uint8_t* or_jpeg_load(const uint8_t* data, size_t len, int req_comps, int* width, int* height,
                      int* actual_comps, float* par, float* dpiY)
{
    if (req_comps != -1 && req_comps != 1 && req_comps != 3 && req_comps != 4) return NULL;
    jd* d = (jd*)calloc(1, sizeof(jd));
    d->data = data; d->len = len;
    /* m_pixelsPerInchX/Y and m_pixelAspectRatio are plain D float members that initit() never assigns:
     * without a JFIF/EXIF density they keep float.init = NaN (jpegload.d:522-524,1971-2074) */
    d->ppiX = d->ppiY = d->par = __builtin_nanf("");
    /* initit (:1971-2074): prime the bit buffer */
    d->bits_left = 16; d->bit_buf = 0;
    get_bits(d, 16); get_bits(d, 16);
    uint8_t* out = NULL;
    /* locate_sof_marker (:1898-1930) */
    if (!locate_soi_marker(d)) goto fail;
    {
        int c = process_markers(d, 0);
        if (c == M_SOF2) d->progressive = 1;                /* locate_sof_marker (:1921-1924) */
        else if (c != M_SOF0 && c != M_SOF1) goto fail;
        if (!read_sof_marker(d)) goto fail;
    }
    *width = d->image_x_size; *height = d->image_y_size; *actual_comps = d->comps_in_frame;
    if (req_comps < 0) req_comps = d->comps_in_frame;
    if (!init_frame(d)) goto fail;
    if (d->progressive) { if (!init_progressive(d)) goto fail; }     /* decode_start (:3696-3706) */
    else {
        if (!init_scan(d)) goto fail;
        if (d->comps_in_scan != d->comps_in_frame) goto fail;   /* non-interleaved multi-scan baseline: unsupported */
    }
    {
        const int W = d->image_x_size, H = d->image_y_size;
        const int dst_bpl = W * req_comps;
        out = (uint8_t*)malloc((size_t)dst_bpl * H);
        for (int y = 0; y < H; ++y) {
            const uint8_t* sl = NULL;
            if (decode_line(d, &sl) != 0) { free(out); out = NULL; goto fail; }
            uint8_t* pDst = out + (size_t)y * dst_bpl;
            const int nc = d->comps_in_frame;
            if ((req_comps == 1 && nc == 1) || (req_comps == 4 && nc == 3)) memcpy(pDst, sl, (size_t)dst_bpl);
            else if (nc == 1) {
                if (req_comps == 3) for (int x = 0; x < W; ++x) { uint8_t l = sl[x]; pDst[0] = l; pDst[1] = l; pDst[2] = l; pDst += 3; }
                else for (int x = 0; x < W; ++x) { uint8_t l = sl[x]; pDst[0] = l; pDst[1] = l; pDst[2] = l; pDst[3] = 255; pDst += 4; }
            } else if (nc == 3) {
                if (req_comps == 1) {
                    const int YR = 19595, YG = 38470, YB = 7471;
                    for (int x = 0; x < W; ++x) { int r = sl[x*4], g = sl[x*4+1], b = sl[x*4+2]; *pDst++ = (uint8_t)((r * YR + g * YG + b * YB + 32768) >> 16); }
                } else for (int x = 0; x < W; ++x) { pDst[0] = sl[x*4]; pDst[1] = sl[x*4+1]; pDst[2] = sl[x*4+2]; pDst += 3; }
            }
        }
        *par = d->par; *dpiY = d->ppiY;
    }
fail:
    jd_free(d);
    free(d);
    return out;
}
