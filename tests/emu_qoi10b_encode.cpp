// emu_qoi10b_encode.cpp -- gamut_b200/csrc/qoi10b_encode.cuh compiled for the host under tests/cuda_emu.h. The launch
// sequence below is the one of gb::qoi10b_encode_device (qoi10b_encode.cu); test infrastructure only.
#include "cuda_emu.h"
#include "../gamut_b200/csrc/qoi10b_encode.cuh"
#include <stdlib.h>
#include <string.h>

struct emu_qoix_desc { uint32_t width, height; int32_t pitchBytes; uint8_t channels, bitdepth, colorspace, compression; float pixelAspectRatio, resolutionY; };

extern "C" int emu_qoi10b_encode_batch(int n, const uint8_t* const* pixels, const emu_qoix_desc* descs, uint8_t* const* outs, int* out_len)
{
    std::vector<QeImage> imgs; std::vector<int> which;
    uint32_t total_tiles = 0, most = 0;
    for (int i = 0; i < n; ++i) {
        out_len[i] = 0;
        const emu_qoix_desc& d = descs[i];
        QeImage Q;
        if (!q10_setup(Q, pixels[i], d.width, d.height, d.pitchBytes, d.channels, d.bitdepth, d.colorspace, d.compression,
                       d.pixelAspectRatio, d.resolutionY, outs[i], total_tiles)) continue;
        imgs.push_back(Q); which.push_back(i);
        most = std::max(most, Q.ntiles);
    }
    const int m = (int)imgs.size();
    if (!m) return 1;
    std::vector<QeTile> tiles((size_t)total_tiles + 1);
    memset(tiles.data(), 0xa5, sizeof(QeTile) * tiles.size());     // device memory is not zeroed either
    std::vector<int> len((size_t)m, -1);
    const dim3 grid(most, (unsigned)m);
    const QeImage* dI = imgs.data(); QeTile* dT = tiles.data(); int* dl = len.data();
    emu::launch(grid, QE_THREADS, [&] { q10_tile_ne_kernel(dI, dT); });
    emu::launch(dim3((unsigned)m), QE_THREADS, [&] { qe_scan_kernel(dI, dT, 0, dl); });
    emu::launch(grid, QE_THREADS, [&] { q10_tile_kernel<false>(dI, dT); });
    emu::launch(dim3((unsigned)m), QE_THREADS, [&] { qe_scan_kernel(dI, dT, 1, dl); });
    emu::launch(grid, QE_THREADS, [&] { q10_tile_kernel<true>(dI, dT); });
    for (int k = 0; k < m; ++k) out_len[which[k]] = len[(size_t)k];
    return 1;
}
