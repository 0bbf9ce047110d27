// qoix_sub.cuh -- the remaining QOIX sub-decoders (included by qoix.cu):
//   qoi2avg_kernel   qoix_decode      (codecs/qoi2avg.d:625-839)  8-bit RGB/RGBA, LOCO-I prediction (:863-897), FIFO index
//   qoiplane8_kernel qoiplane_decode  (codecs/qoiplane.d:377-541) 8-bit L/LA, nibble-aligned opcodes
//   qoi10b_kernel    qoi10b_decode    (codecs/qoi10b.d:504-869)   10-bit 1-4 channels, 2-bit aligned opcodes
// One thread per image in this round (the FIFO index of QOI2AVG and the run/row coupling make the pixel
// recurrence serial in raster order; the parse could be chunk-parallel like QOI-Plane10 -- next round).
// Reads past the end of a stream yield 0xFF, memory the reference leaves uninitialised is zero.
#pragma once

struct SubJob {
    const uint8_t* stream; uint32_t size;
    uint8_t* out; uint32_t w, h;
    int channels, version, image;
    uint8_t* rows;              // scratch: two zeroed scanlines (RGBA8 or 4 x u16 per pixel)
};

struct ByteSrcD {
    const uint8_t* b; uint32_t size;
    __device__ __forceinline__ uint32_t at(uint32_t p) const { return p < size ? b[p] : 0xFFu; }
};

__device__ __forceinline__ int loco8(int a, int b, int c)          // qoi2avg.d:863-897
{
    const int mx = max(a, b), mn = min(a, b);
    int p = a + b - c;
    if (c >= mx) p = mn;
    if (c <= mn) p = mx;
    return min(max(p, 0), 255);
}

__global__ void __launch_bounds__(32)
qoi2avg_kernel(const SubJob* __restrict__ jobs, int njobs, const int* __restrict__ status)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= njobs) return;
    const SubJob J = jobs[j];
    if (!status[J.image]) return;
    const ByteSrcD S{J.stream, J.size};
    const int W = (int)J.w, H = (int)J.h, ch = J.channels;
    uchar4* cur = (uchar4*)J.rows; uchar4* last = cur + W;
    uchar4 index[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) index[i] = make_uchar4(0, 0, 0, 0);
    uchar4 px = make_uchar4(0, 0, 0, 255), ref;
    uint32_t p = 25; int run = 0, index_pos = 0;
    const uint32_t chunks_len = J.size - 4;
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            if (run > 0) --run;
            else if (p < chunks_len) {
                ref = px;
                if (y > 0) {
                    const uchar4 up = last[x];
                    if (x == 0) { ref.x = up.x; ref.y = up.y; ref.z = up.z; }
                    else {
                        const uchar4 ul = last[x - 1];
                        ref.x = (uint8_t)loco8(px.x, up.x, ul.x); ref.y = (uint8_t)loco8(px.y, up.y, ul.y); ref.z = (uint8_t)loco8(px.z, up.z, ul.z);
                    }
                }
                bool end = false;
                for (;;) {
                    const int b1 = (int)S.at(p++);
                    if (b1 < 0x80) {
                        const int vg = ((b1 >> 4) & 7) - 4, bias = vg < 0 ? 1 : 2;
                        px.y = (uint8_t)(ref.y + vg);
                        px.x = (uint8_t)(ref.x + vg - bias + ((b1 >> 2) & 3));
                        px.z = (uint8_t)(ref.z + vg - bias + (b1 & 3));
                        index[index_pos++ & 63] = px;
                    } else if (b1 < 0xc0) px = index[b1 & 63];
                    else if (b1 < 0xe0) {
                        const int b2 = (int)S.at(p++), vg = (b1 & 0x1f) - 16;
                        px.x = (uint8_t)(ref.x + vg - 8 + ((b2 >> 4) & 15)); px.y = (uint8_t)(ref.y + vg); px.z = (uint8_t)(ref.z + vg - 8 + (b2 & 15));
                        index[index_pos++ & 63] = px;
                    } else if (b1 < 0xe8) {
                        int dv = (b1 << 8) | (int)S.at(p++); dv = (dv << 8) | (int)S.at(p++);
                        const int vg = ((dv >> 12) & 0x7f) - 64;
                        px.x = (uint8_t)(ref.x + vg + ((dv >> 6) & 0x3f) - 32); px.y = (uint8_t)(ref.y + vg); px.z = (uint8_t)(ref.z + vg + (dv & 0x3f) - 32);
                        index[index_pos++ & 63] = px;
                    } else if (b1 < 0xf0) { px.w = (uint8_t)(px.w + (b1 & 7) - 4); continue; }
                    else if (b1 < 0xf8) run = b1 & 7;
                    else if (b1 < 0xfc) run = ((b1 & 3) << 8) | (int)S.at(p++);
                    else if (b1 == 0xfc) { const uint8_t v = (uint8_t)S.at(p++); px.x = px.y = px.z = v; index[index_pos++ & 63] = px; }
                    else if (b1 == 0xfd) { px.x = (uint8_t)S.at(p++); px.y = (uint8_t)S.at(p++); px.z = (uint8_t)S.at(p++); index[index_pos++ & 63] = px; }
                    else if (b1 == 0xfe) { px.x = (uint8_t)S.at(p++); px.y = (uint8_t)S.at(p++); px.z = (uint8_t)S.at(p++); px.w = (uint8_t)S.at(p++); index[index_pos++ & 63] = px; }
                    else end = true;
                    break;
                }
                if (end) break;
            }
            cur[x] = px;
        }
        uint8_t* line = J.out + (size_t)y * W * ch;
        if (ch == 4) { for (int x = 0; x < W; ++x) ((uchar4*)line)[x] = cur[x]; }
        else for (int x = 0; x < W; ++x) { const uchar4 q = cur[x]; line[x * 3] = q.x; line[x * 3 + 1] = q.y; line[x * 3 + 2] = q.z; }
        uchar4* t = cur; cur = last; last = t;
    }
}

__global__ void __launch_bounds__(32)
qoiplane8_kernel(const SubJob* __restrict__ jobs, int njobs, const int* __restrict__ status)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= njobs) return;
    const SubJob J = jobs[j];
    if (!status[J.image]) return;
    const ByteSrcD S{J.stream, J.size};
    const int W = (int)J.w, H = (int)J.h, ch = J.channels;
    uint32_t np = 50;     // nibble position (25 bytes of header)
    auto nib = [&]() -> int { const uint32_t b = S.at(np >> 1); const int v = (np & 1) ? (int)(b & 15) : (int)(b >> 4); ++np; return v; };
    auto ubyte = [&]() -> int { const int hi = nib() << 4; return hi | nib(); };
    int l = 0, a = 255, run = 0;
    for (int y = 0; y < H; ++y) {
        uint8_t* line = J.out + (size_t)y * W * ch;
        const uint8_t* above = y ? line - (size_t)W * ch : nullptr;
        for (int x = 0; x < W; ++x) {
            const int ref_l = l, ref_a = a;
            if (run > 0) --run;
            else {
                for (;;) {
                    const int op = nib();
                    if (op == 0xf) { run = ubyte() + 3; if (run == 258) run = 0x7fffffff; }
                    else if ((op & 0xc) == 0xc) run = op & 3;
                    else {
                        const int top = y ? above[x * ch] : ref_l;
                        const int avg = (top + ref_l + 1) >> 1;
                        if ((op & 8) == 0) l = (avg + op - 4) & 255;
                        else if ((op & 0xe) == 8) { const int v = ((op & 1) << 4) + nib(); l = (avg + v - 16) & 255; }
                        else if (op == 0xa) l = ubyte();
                        else {
                            const int diff = nib();
                            if (diff == 0) { l = ubyte(); a = ubyte(); }
                            else { a = (ref_a + diff - 8) & 255; continue; }
                        }
                    }
                    break;
                }
            }
            if (ch == 1) line[x] = (uint8_t)l;
            else { line[x * 2] = (uint8_t)l; line[x * 2 + 1] = (uint8_t)a; }
        }
    }
}

__device__ __forceinline__ int loco10(int a, int b, int c)         // qoi10b.d:871-903
{
    const int mx = max(a, b), mn = min(a, b);
    int p = a + b - c;
    if (c >= mx) p = mn;
    if (c <= mn) p = mx;
    return min(max(p, 0), 1023);
}

__global__ void __launch_bounds__(32)
qoi10b_kernel(const SubJob* __restrict__ jobs, int njobs, const int* __restrict__ status)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= njobs) return;
    const SubJob J = jobs[j];
    if (!status[J.image]) return;
    const ByteSrcD S{J.stream, J.size};
    const int W = (int)J.w, H = (int)J.h, ch = J.channels;
    ushort4* cur = (ushort4*)J.rows; ushort4* last = cur + W;
    uint32_t bp = 25 * 8;     // bit position, MSB first
    auto bits = [&](int n) -> uint32_t {      // n <= 24
        const uint32_t byte0 = bp >> 3;
        const uint32_t v = (S.at(byte0) << 24) | (S.at(byte0 + 1) << 16) | (S.at(byte0 + 2) << 8) | S.at(byte0 + 3);
        const uint32_t r = (v << (bp & 7)) >> (32 - n);
        bp += n;
        return r;
    };
    auto sx = [](uint32_t v, int b) -> int { return (int)(v << (32 - b)) >> (32 - b); };
    const bool grey = ch <= 2;
    ushort4 px = make_ushort4(0, 0, 0, 1023), ref;
    int run = 0; bool finished = false;
    uint16_t* out16 = (uint16_t*)J.out;
    int y = 0;
    for (; y < H && !finished; ++y) {
        for (int x = 0; x < W; ++x) {
            ref = px;
            if (run > 0) --run;
            else {
                if (y > 0) {
                    const ushort4 up = last[x];
                    if (J.version >= 2) {
                        if (x == 0) { ref.x = up.x; ref.y = up.y; ref.z = up.z; }
                        else {
                            const ushort4 ul = last[x - 1];
                            ref.x = (uint16_t)loco10(px.x, up.x, ul.x); ref.y = (uint16_t)loco10(px.y, up.y, ul.y); ref.z = (uint16_t)loco10(px.z, up.z, ul.z);
                        }
                    } else {
                        ref.x = (uint16_t)((ref.x + up.x + 1) >> 1); ref.y = (uint16_t)((ref.y + up.y + 1) >> 1); ref.z = (uint16_t)((ref.z + up.z + 1) >> 1);
                    }
                }
                for (;;) {
                    const int op = (int)bits(8);
                    if (op < 0x80) {
                        const int vg = sx((op >> 2) & 31, 5);
                        px.y = (uint16_t)((ref.y + vg) & 1023);
                        if (!grey) {
                            const int vg_r = sx(((op & 3) << 2) | bits(2), 4), vg_b = sx(bits(4), 4);
                            px.x = (uint16_t)((ref.x + vg + vg_r) & 1023); px.z = (uint16_t)((ref.z + vg + vg_b) & 1023);
                        } else { bp -= 2; px.x = px.y; px.z = px.y; }
                    } else if (op < 0xc0) {
                        const int vg = sx((op >> 2) & 15, 4);
                        px.y = (uint16_t)((ref.y + vg) & 1023);
                        if (!grey) {
                            const uint32_t remain = bits(4);
                            const int vg_r = sx(((op & 3) << 1) | (remain >> 3), 3), vg_b = sx(remain & 7, 3);
                            px.x = (uint16_t)((ref.x + vg + vg_r) & 1023); px.z = (uint16_t)((ref.z + vg + vg_b) & 1023);
                        } else { bp -= 2; px.x = px.y; px.z = px.y; }
                    } else if (op < 0xe0) {
                        const int vg = sx(((op & 31) << 2) | bits(2), 7);
                        px.y = (uint16_t)((ref.y + vg) & 1023);
                        if (!grey) {
                            const int vg_r = sx(bits(6), 6), vg_b = sx(bits(6), 6);
                            px.x = (uint16_t)((ref.x + vg + vg_r) & 1023); px.z = (uint16_t)((ref.z + vg + vg_b) & 1023);
                        } else { px.x = px.y; px.z = px.y; }
                    } else if (op < 0xe8) {
                        const int vg = sx(((op & 7) << 6) | bits(6), 9);
                        px.y = (uint16_t)((ref.y + vg) & 1023);
                        if (!grey) {
                            const int vg_r = sx(bits(8), 8), vg_b = sx(bits(8), 8);
                            px.x = (uint16_t)((ref.x + vg + vg_r) & 1023); px.z = (uint16_t)((ref.z + vg + vg_b) & 1023);
                        } else { px.x = px.y; px.z = px.y; }
                    } else if (op < 0xf0) { px.w = (uint16_t)((px.w + sx(((op & 7) << 2) | bits(2), 5)) & 1023); continue; }
                    else if ((op & 0xfc) == 0xf8) { px.w = (uint16_t)((px.w + sx(((op & 3) << 6) | bits(6), 8)) & 1023); continue; }
                    else if (op < 0xf8) { run = op & 7; if (run == 7) run = (int)bits(8) + 7; }
                    else if (op == 0xfd || op == 0xfe) {
                        px.x = (uint16_t)bits(10);
                        if (!grey) { px.y = (uint16_t)bits(10); px.z = (uint16_t)bits(10); } else { px.y = px.x; px.z = px.x; }
                        if (op == 0xfe) px.w = (uint16_t)bits(10);
                    } else if (op == 0xfc) { px.x = (uint16_t)bits(10); px.y = px.x; px.z = px.x; }
                    else finished = true;
                    break;
                }
                if (finished) break;
            }
            cur[x] = px;
        }
        if (finished) break;
        uint16_t* line = out16 + (size_t)y * W * ch;
        for (int x = 0; x < W; ++x) {
            const ushort4 q = cur[x];
            const uint16_t r = (uint16_t)(q.x << 6 | (q.x >> 4)), g = (uint16_t)(q.y << 6 | (q.y >> 4));
            const uint16_t b = (uint16_t)(q.z << 6 | (q.z >> 4)), a = (uint16_t)(q.w << 6 | (q.w >> 4));
            if (ch == 4) { line[x * 4] = r; line[x * 4 + 1] = g; line[x * 4 + 2] = b; line[x * 4 + 3] = a; }
            else if (ch == 3) { line[x * 3] = r; line[x * 3 + 1] = g; line[x * 3 + 2] = b; }
            else if (ch == 2) { line[x * 2] = r; line[x * 2 + 1] = a; }
            else line[x] = r;
        }
        ushort4* t = cur; cur = last; last = t;
    }
    // rows from the one that met END onwards are never written by the reference: zero here
    for (size_t i = (size_t)y * W * ch; i < (size_t)H * W * ch; ++i) out16[i] = 0;
}
