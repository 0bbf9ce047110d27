// emu_tga.cpp -- gamut_b200/csrc/tga.cuh (header walk + the two TGA kernels) compiled for the host under
// tests/cuda_emu.h. The launch sequence below is the one of gb::tga_decode_batch (tga.cu); test infrastructure only.
#include "cuda_emu.h"
#include "../gamut_b200/csrc/tga.cuh"
#include <stdlib.h>

// Decodes one file into out (capacity out_cap); returns 1 and w / h / comp, or 0 where the decoder fails.
extern "C" int emu_tga_load(const uint8_t* data, size_t len, uint8_t* out, size_t out_cap, int* w, int* h, int* comp)
{
    TgaPlan P;
    if (len > 0xfffffff0u || !tga_plan(data, len, P)) return 0;
    if ((size_t)P.w * P.h * P.components > out_cap) return -1;
    int fail = 0;
    TgaJob J; memset(&J, 0, sizeof(J));
    J.data = data; J.palette = P.palette.empty() ? nullptr : P.palette.data(); J.out = out; J.fail = &fail;
    J.len = (uint32_t)len; J.pix_off = P.pix_off; J.palette_len = P.palette_len; J.pix_base = 0;
    J.w = P.w; J.h = P.h; J.components = P.components; J.src_bytes = P.src_bytes; J.mode = P.mode; J.index16 = P.index16;
    J.inverted = P.inverted; J.rle = P.rle;
    const TgaJob* dj = &J;
    const uint32_t total = (uint32_t)P.w * (uint32_t)P.h;
    if (!P.rle) emu::launch(dim3((total + 255) / 256), 256, [&] { tga_raw_kernel(dj, 1, total); });
    else {
        const uint32_t ms = tga_max_segments(total);
        std::vector<TgaCheckpoint> ck((size_t)ms + 1, TgaCheckpoint{0xdeadbeefu, 0xdeadbeefu});
        uint32_t nseg = 0xdeadbeefu;
        TgaCheckpoint* dc = ck.data(); uint32_t* dn = &nseg;
        emu::launch(dim3(1), 32, [&] { tga_rle_index_kernel(dj, dc, dn); });
        if (!fail && nseg > ms) return -2;
        emu::launch(dim3(ms, 1), 32, [&] { tga_rle_kernel(dj, dc, dn); });
    }
    if (fail) return 0;
    *w = P.w; *h = P.h; *comp = P.components;
    return 1;
}

// ---- encoder: the launches of gb::tga_encode_device (tga.cu) -------------------------------------------------------------
#include "../gamut_b200/csrc/tga_encode.cuh"

extern "C" int emu_tga_encode(const uint8_t* pixels, int type, int width, int height, int pitch, uint8_t* out, size_t out_cap)
{
    TeImage T; uint32_t total_rows = 0;
    if (te_bound(type, width, height) > out_cap) return -1;
    if (!te_setup(T, pixels, type, width, height, pitch, out, total_rows)) return 0;
    std::vector<uint32_t> bytes((size_t)total_rows + 1, 0xdeadbeefu), off((size_t)total_rows + 1, 0xdeadbeefu);
    int len = -1;
    const TeImage* dI = &T; uint32_t* dB = bytes.data(); uint32_t* dO = off.data(); int* dl = &len;
    const dim3 grid((unsigned)T.h, 1);
    emu::launch(grid, 32, [&] { te_row_kernel<false>(dI, dB, dO); });
    emu::launch(dim3(1), 256, [&] { te_scan_kernel(dI, dB, dO, dl); });
    emu::launch(grid, 32, [&] { te_row_kernel<true>(dI, dB, dO); });
    return len;
}
