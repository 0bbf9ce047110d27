// tga.cuh -- the header walk and the kernels of tga.cu (TGA decoder). Kept apart from the CUDA host code
// so that tests/test_tga_emulated.py can compile exactly this text for the host under a thread-per-CUDA-thread
// emulation (tests/cuda_emu.h) and compare it with the oracle without a GPU.
//
// Reference: TGADecoder.getImageInfo / decodeImage (codecs/tga.d:313-588) as loadTGA calls them (plugins/tga.d:45-105),
// reading from a MemoryFile (io.d:384-440: a read past the end fails, a seek may land on the end).
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <vector>

namespace {

enum { TGA_RAW = 0, TGA_RGB16 = 1, TGA_INDEXED = 2 };

struct TgaJob {
    const uint8_t* data;        // the file
    const uint8_t* palette;     // palette_len x components bytes in output channel order (TGA_INDEXED)
    uint8_t* out;               // w x h x components, gapless
    int* fail;                  // set to 1 when a read of the packet walk fails
    uint32_t len, pix_off;      // file length; offset of the first pixel / packet
    uint32_t palette_len, pix_base;
    uint32_t ck_base;            // first checkpoint of this image (run-length files)
    int w, h, components, src_bytes, mode, index16, inverted, rle;
};

// what the header says (getImageInfo, :313-382, and the first lines of decodeImage, :384-420); `palette` is filled in
// output channel order: 15/16-bit entries through stbi__tga_read_rgb16 (:619-646), 24/32-bit entries with B and R
// swapped (the swap decodeImage does over the finished image, :553-565, moved to the table)
struct TgaPlan {
    bool ok = false;
    int w = 0, h = 0, components = 0, src_bytes = 0, mode = 0, index16 = 0, inverted = 0, rle = 0;
    uint32_t pix_off = 0, palette_len = 0;
    std::vector<uint8_t> palette;
};

inline void tga_rgb16(uint32_t px, uint8_t* out)                 // stbi__tga_read_rgb16 (:619-646)
{
    out[0] = (uint8_t)((((px >> 10) & 31u) * 255u) / 31u);
    out[1] = (uint8_t)((((px >> 5) & 31u) * 255u) / 31u);
    out[2] = (uint8_t)(((px & 31u) * 255u) / 31u);
}

inline int tga_get_comp(int bits, bool is_grey, bool* rgb16)     // stbi__tga_get_comp (:590-617)
{
    *rgb16 = false;
    switch (bits) {
    case 8: return 1;
    case 16: if (is_grey) return 2; *rgb16 = true; return 3;
    case 15: *rgb16 = true; return 3;
    case 24: case 32: return bits / 8;
    default: return 0;
    }
}

inline bool tga_plan(const uint8_t* d, size_t len, TgaPlan& P)
{
    size_t p = 0;
    auto r8 = [&](int& v) { if (!d || p + 1 > len) return false; v = d[p++]; return true; };
    auto r16 = [&](int& v) { if (!d || p + 2 > len) return false; v = d[p] | (d[p + 1] << 8); p += 2; return true; };
    auto sk = [&](size_t n) { if (p + n > len) return false; p += n; return true; };
    int idlen = 0, cmap = 0, type = 0, pal_start = 0, pal_len = 0, cmap_bits = 0, w = 0, h = 0, bpp = 0, descriptor = 0;
    if (!r8(idlen) || !r8(cmap) || cmap > 1 || !r8(type)) return false;
    if (cmap == 1) {
        if (type != 1 && type != 9) return false;
        if (!r16(pal_start) || !r16(pal_len) || pal_len == 0 || !r8(cmap_bits)) return false;
        if (cmap_bits != 8 && cmap_bits != 15 && cmap_bits != 16 && cmap_bits != 24 && cmap_bits != 32) return false;
        if (!sk(4)) return false;
    } else {
        if (type != 2 && type != 3 && type != 10 && type != 11) return false;
        if (!sk(9)) return false;
    }
    if (!r16(w) || !r16(h) || w < 1 || h < 1 || !r8(bpp)) return false;
    if (cmap == 1 && bpp != 8 && bpp != 16) return false;
    if (bpp != 8 && bpp != 15 && bpp != 16 && bpp != 24 && bpp != 32) return false;
    // decodeImage
    if (type >= 8) { type -= 8; P.rle = 1; }
    if (!r8(descriptor)) return false;
    P.inverted = 1 - ((descriptor >> 5) & 1);
    bool rgb16 = false;
    P.components = cmap ? tga_get_comp(cmap_bits, false, &rgb16) : tga_get_comp(bpp, type == 3, &rgb16);
    if (!P.components || !sk((size_t)idlen)) return false;
    P.w = w; P.h = h;
    if (cmap) {
        if (!sk((size_t)pal_start)) return false;
        P.mode = TGA_INDEXED; P.index16 = bpp == 16; P.src_bytes = bpp == 16 ? 2 : 1; P.palette_len = (uint32_t)pal_len;
        P.palette.resize((size_t)pal_len * P.components);
        if (rgb16) {
            if (p + (size_t)pal_len * 2 > len) return false;
            for (int i = 0; i < pal_len; ++i, p += 2) tga_rgb16((uint32_t)(d[p] | d[p + 1] << 8), &P.palette[(size_t)i * 3]);
        } else {
            const size_t bytes = (size_t)pal_len * P.components;
            if (p + bytes > len) return false;
            memcpy(P.palette.data(), d + p, bytes);
            p += bytes;
            if (P.components >= 3)
                for (int i = 0; i < pal_len; ++i) { uint8_t* e = &P.palette[(size_t)i * P.components]; const uint8_t t = e[0]; e[0] = e[2]; e[2] = t; }
        }
    } else if (rgb16) { P.mode = TGA_RGB16; P.src_bytes = 2; }
    else { P.mode = TGA_RAW; P.src_bytes = P.components; }
    P.pix_off = (uint32_t)p;
    // without packets every pixel is read: a file that is too short fails (row reads :423-434, pixel reads :488-527)
    if (!P.rle && p + (size_t)w * h * P.src_bytes > len) return false;
    P.ok = true;
    return true;
}

// one source pixel -> components bytes (:488-533; the B/R swap of :553-565 for 24/32-bit pixels happens here)
__device__ __forceinline__ void tga_pixel(const TgaJob& J, const uint8_t* src, uint8_t* dst)
{
    if (J.mode == TGA_INDEXED) {
        uint32_t idx = J.index16 ? (uint32_t)src[0] | (uint32_t)src[1] << 8 : src[0];
        if (idx >= J.palette_len) idx = 0;                        // invalid index (:499-503)
        const uint8_t* e = J.palette + (size_t)idx * J.components;
        for (int j = 0; j < J.components; ++j) dst[j] = e[j];
    } else if (J.mode == TGA_RGB16) {
        const uint32_t px = (uint32_t)src[0] | (uint32_t)src[1] << 8;
        dst[0] = (uint8_t)((((px >> 10) & 31u) * 255u) / 31u);
        dst[1] = (uint8_t)((((px >> 5) & 31u) * 255u) / 31u);
        dst[2] = (uint8_t)(((px & 31u) * 255u) / 31u);
    } else if (J.components >= 3) {
        dst[0] = src[2]; dst[1] = src[1]; dst[2] = src[0];
        if (J.components == 4) dst[3] = src[3];
    } else {
        dst[0] = src[0];
        if (J.components == 2) dst[1] = src[1];
    }
}
// pixel i of the stream lands in row i / w counted from the bottom unless the descriptor says top-down (:395, :537-551)
__device__ __forceinline__ uint8_t* tga_dest(const TgaJob& J, uint32_t i)
{
    const uint32_t row = i / (uint32_t)J.w, x = i - row * (uint32_t)J.w;
    const uint32_t r = J.inverted ? (uint32_t)J.h - 1u - row : row;
    return J.out + ((size_t)r * J.w + x) * J.components;
}

// ---- T1: files without packets, one thread per pixel over all such images of the batch ---------------------------------
__global__ void __launch_bounds__(256)
tga_raw_kernel(const TgaJob* __restrict__ jobs, int njobs, uint32_t total)
{
    const uint32_t g = blockIdx.x * 256u + threadIdx.x;
    if (g >= total) return;
    int lo = 0, hi = njobs - 1;                                   // the job whose pixel range holds g
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (jobs[mid].pix_base <= g) lo = mid; else hi = mid - 1; }
    const TgaJob& J = jobs[lo];
    const uint32_t i = g - J.pix_base;
    tga_pixel(J, J.data + J.pix_off + (size_t)i * J.src_bytes, tga_dest(J, i));
}

// ---- T2 / T3: run-length packets (:468-486, :535). The packet chain is serial by the format (a packet header says
// where the next one is) and a packet may cross rows, so nothing tells where a row starts. Two passes:
//   T2 tga_rle_index_kernel: one warp per image walks the headers only -- no pixel is touched -- and writes a checkpoint
//      (file offset, pixel index) every TGA_SEG packets; a read past the end of the file fails the image here.
//   T3 tga_rle_kernel: one warp per checkpoint walks its TGA_SEG packets and its lanes place the pixels (at most 128 per
//      packet): thousands of independent warps per image instead of one.
// Both stage the stream through an 8 KB window in shared memory (coalesced word loads), so a header read is a shared-
// memory access, not a dependent global load. A packet that runs past the last pixel is cut (:463, :535).
constexpr uint32_t TGA_WIN = 8192, TGA_PACKET_MAX = 1 + 128 * 4, TGA_SEG = 256;

struct TgaCheckpoint { uint32_t pos, pixel; };
// checkpoints of job j: ck[J.ck_base .. J.ck_base + tga_max_segments(J)); the number in use is nseg[j]
__host__ __device__ inline uint32_t tga_max_segments(uint32_t total_pixels) { return total_pixels / TGA_SEG + 1u; }   // a packet holds >= 1 pixel

__device__ __forceinline__ void tga_refill(const TgaJob& J, uint8_t* s_win, uint32_t pos, uint32_t& win_start, uint32_t& win_end, uint32_t lane)
{
    // from a 4-byte aligned address at or below pos (pos >= 18, so this never reaches before the file)
    __syncwarp();
    win_start = pos - (uint32_t)((uintptr_t)(J.data + pos) & 3u);
    win_end = min(win_start + TGA_WIN, J.len);
    const uint32_t nbytes = win_end - win_start, nwords = nbytes >> 2;
    const uint32_t* src = (const uint32_t*)(J.data + win_start);
    for (uint32_t k = lane; k < nwords; k += 32u) ((uint32_t*)s_win)[k] = src[k];
    for (uint32_t k = (nwords << 2) + lane; k < nbytes; k += 32u) s_win[k] = J.data[win_start + k];
    __syncwarp();
}

__global__ void __launch_bounds__(32)
tga_rle_index_kernel(const TgaJob* __restrict__ jobs, TgaCheckpoint* __restrict__ ck, uint32_t* __restrict__ nseg)
{
    __shared__ __align__(16) uint8_t s_win[TGA_WIN];
    const TgaJob& J = jobs[blockIdx.x];
    const uint32_t total = (uint32_t)J.w * (uint32_t)J.h, lane = threadIdx.x;
    TgaCheckpoint* C = ck + J.ck_base;
    uint32_t pos = J.pix_off, i = 0, packets = 0, segs = 0;
    uint32_t win_start = 0, win_end = 0;                          // file offsets held in s_win
    while (i < total) {                                           // every lane walks the same chain
        if (pos + 2u > win_end && win_end < J.len) tga_refill(J, s_win, pos, win_start, win_end, lane);
        if ((packets & (TGA_SEG - 1u)) == 0u) { if (lane == 0) { C[segs].pos = pos; C[segs].pixel = i; } ++segs; }
        if (pos + 1u > J.len) { if (lane == 0) *J.fail = 1; break; }
        const uint32_t cmd = s_win[pos - win_start];
        const uint32_t count = 1u + (cmd & 127u), rep = cmd >> 7;
        const uint32_t n = min(count, total - i);
        const uint32_t need = rep ? (uint32_t)J.src_bytes : n * (uint32_t)J.src_bytes;
        if ((unsigned long long)pos + 1u + need > J.len) { if (lane == 0) *J.fail = 1; break; }
        pos += 1u + (rep ? (uint32_t)J.src_bytes : count * (uint32_t)J.src_bytes);
        i += n;
        ++packets;
    }
    if (lane == 0) nseg[blockIdx.x] = segs;
}

// grid = (most segments of an image, images); the jobs' fail flags are already final (same stream, after T2)
__global__ void __launch_bounds__(32)
tga_rle_kernel(const TgaJob* __restrict__ jobs, const TgaCheckpoint* __restrict__ ck, const uint32_t* __restrict__ nseg)
{
    __shared__ __align__(16) uint8_t s_win[TGA_WIN];
    const TgaJob& J = jobs[blockIdx.y];
    if (blockIdx.x >= nseg[blockIdx.y] || *J.fail) return;
    const uint32_t total = (uint32_t)J.w * (uint32_t)J.h, lane = threadIdx.x;
    const TgaCheckpoint c = ck[J.ck_base + blockIdx.x];
    uint32_t pos = c.pos, i = c.pixel;
    uint32_t win_start = 0, win_end = 0;
    for (uint32_t packets = 0; packets < TGA_SEG && i < total; ++packets) {
        if (pos + TGA_PACKET_MAX > win_end && win_end < J.len) tga_refill(J, s_win, pos, win_start, win_end, lane);
        const uint32_t cmd = s_win[pos - win_start];              // T2 has checked every read of this walk
        const uint32_t count = 1u + (cmd & 127u), rep = cmd >> 7;
        const uint32_t n = min(count, total - i);
        const uint8_t* pk = s_win + (pos - win_start) + 1u;
        for (uint32_t k = lane; k < n; k += 32u)
            tga_pixel(J, pk + (rep ? 0u : k * (uint32_t)J.src_bytes), tga_dest(J, i + k));
        pos += 1u + (rep ? (uint32_t)J.src_bytes : count * (uint32_t)J.src_bytes);
        i += n;
    }
}

}  // namespace
