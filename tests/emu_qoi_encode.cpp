// emu_qoi_encode.cpp -- gamut_b200/csrc/qoi_encode.cuh compiled for the host under tests/cuda_emu.h. The launch
// sequence below is the one of gb::qoi_encode_device (qoi_encode.cu); test infrastructure only.
#include "cuda_emu.h"
#include "../gamut_b200/csrc/qoi_encode.cuh"
#include <stdlib.h>
#include <string.h>

extern "C" int emu_qoi_encode_batch(int n, const uint8_t* const* pixels, const uint32_t* widths, const uint32_t* heights,
                                    const int* channels, const int* colorspaces, const int* pitches, uint8_t* const* outs,
                                    int* out_len)
{
    std::vector<QnImage> imgs; std::vector<int> which;
    uint32_t total_tiles = 0;
    for (int i = 0; i < n; ++i) {
        out_len[i] = 0;
        QnImage Q;
        if (!qn_setup(Q, pixels[i], widths[i], heights[i], channels[i], colorspaces[i], pitches[i], outs[i], total_tiles)) continue;
        imgs.push_back(Q); which.push_back(i);
    }
    const int m = (int)imgs.size();
    if (!m) return 1;
    std::vector<QnTile> tiles((size_t)total_tiles + 1);
    std::vector<uint32_t> val(64 * ((size_t)total_tiles + 1), 0xdeadbeefu);
    std::vector<int> len((size_t)m, -1);
    memset(tiles.data(), 0xa5, sizeof(QnTile) * tiles.size());     // device memory is not zeroed either
    uint32_t most = 0;
    for (const QnImage& Q : imgs) most = std::max(most, Q.ntiles);
    const dim3 grid(most, (unsigned)m);
    const QnImage* dI = imgs.data(); QnTile* dT = tiles.data(); uint32_t* dV = val.data(); int* dl = len.data();
    emu::launch(grid, QN_THREADS, [&] { qn_tile_state_kernel(dI, dT, dV); });
    emu::launch(dim3((unsigned)m), QN_THREADS, [&] { qn_scan_kernel(dI, dT, dV, 0, dl); });
    emu::launch(grid, QN_THREADS, [&] { qn_tile_kernel<false>(dI, dT, dV); });
    emu::launch(dim3((unsigned)m), QN_THREADS, [&] { qn_scan_kernel(dI, dT, dV, 1, dl); });
    emu::launch(grid, QN_THREADS, [&] { qn_tile_kernel<true>(dI, dT, dV); });
    for (int k = 0; k < m; ++k) out_len[which[k]] = len[(size_t)k];
    return 1;
}
