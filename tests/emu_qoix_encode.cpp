// emu_qoix_encode.cpp -- gamut_b200/csrc/qoix_encode.cuh (QOI-Plane10 and QOI-Plane encoders) compiled for the host
// under tests/cuda_emu.h. The launch sequence below is the one of gb::qoiplane_encode_device (qoix_encode.cu); test
// infrastructure only.
#include "cuda_emu.h"
#include "../gamut_b200/csrc/qoix_encode.cuh"
#include <stdlib.h>
#include <string.h>

struct emu_qoix_desc { uint32_t width, height; int32_t pitchBytes; uint8_t channels, bitdepth, colorspace, compression; float pixelAspectRatio, resolutionY; };

template <bool P8>
static void run(const QeImage* dI, int m, uint32_t most, QeTile* dT, int* dl)
{
    const dim3 grid(most, (unsigned)m);
    emu::launch(grid, QE_THREADS, [&] { qe_tile_ne_kernel<P8>(dI, m, dT); });
    emu::launch(dim3((unsigned)m), QE_THREADS, [&] { qe_scan_kernel(dI, dT, 0, dl); });
    emu::launch(grid, QE_THREADS, [&] { qe_tile_kernel<false, P8>(dI, m, dT); });
    emu::launch(dim3((unsigned)m), QE_THREADS, [&] { qe_scan_kernel(dI, dT, 1, dl); });
    emu::launch(grid, QE_THREADS, [&] { qe_tile_kernel<true, P8>(dI, m, dT); });
}

extern "C" int emu_qoix_encode_batch(int n, const uint8_t* const* pixels, const emu_qoix_desc* descs, uint8_t* const* outs, int* out_len)
{
    std::vector<QeImage> imgs; std::vector<int> which;
    uint32_t total_tiles = 0, most[2] = {0, 0};
    int count[2] = {0, 0};
    for (int i = 0; i < n; ++i) out_len[i] = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (int i = 0; i < n; ++i) {
            const emu_qoix_desc& d = descs[i];
            if ((d.bitdepth == 8) != (pass == 1)) continue;
            QeImage Q;
            if (!qe_setup(Q, pixels[i], d.width, d.height, d.pitchBytes, d.channels, d.bitdepth, d.colorspace, d.compression,
                          d.pixelAspectRatio, d.resolutionY, outs[i], total_tiles)) continue;
            imgs.push_back(Q); which.push_back(i);
            ++count[pass]; most[pass] = std::max(most[pass], Q.ntiles);
        }
    const int m = (int)imgs.size();
    if (!m) return 1;
    std::vector<QeTile> tiles((size_t)total_tiles + 1);
    memset(tiles.data(), 0xa5, sizeof(QeTile) * tiles.size());     // device memory is not zeroed either
    std::vector<int> len((size_t)m, -1);
    if (count[0]) run<false>(imgs.data(), count[0], most[0], tiles.data(), len.data());
    if (count[1]) run<true>(imgs.data() + count[0], count[1], most[1], tiles.data(), len.data() + count[0]);
    for (int k = 0; k < m; ++k) out_len[which[k]] = len[(size_t)k];
    return 1;
}
