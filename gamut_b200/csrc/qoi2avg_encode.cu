// qoi2avg_encode.cu -- QOI2AVG encoder on the GPU: qoix_encode (codecs/qoi2avg.d:376-617), the codec
// qoix_lz4_encode / saveQOIX pick for rgb8 / rgba8 / rgbap8 images (plugins/qoix.d:184-200, :251-339). Kernels and their
// description: qoi2avg_encode.cuh. Reached through gb200_qoix_encode / gb200_qoix_encode_batch_device (qoix_encode.cu)
// for descs with 3 or 4 channels. As for the other QOIX sub-encoders the LZ4 stage is not built: compression = 0.
#include "../../include/gamut_b200.h"
#include "common.h"
#include "qoi2avg_encode.cuh"
#include <algorithm>
#include <vector>
#include <cstring>

namespace gb {

// Encodes the 3 / 4-channel 8-bit images of a batch (other entries are left alone: out_len[i] is only written for
// those). out_dev[i]: at least gb200_qoix_encode_bound bytes, 16-byte aligned.
bool qoi2avg_encode_device(int n, const uint8_t* const* pixels_dev, const gb200_qoix_desc* descs, uint8_t* const* out_dev,
                           int* out_len, cudaStream_t st)
{
    if (!ensure_device()) return false;
    size_t ix_total = 0;
    for (int i = 0; i < n; ++i)
        if (descs[i].channels >= 3 && q2_valid(descs[i].width, descs[i].height, descs[i].channels, descs[i].bitdepth, descs[i].colorspace, descs[i].compression))
            ix_total += ((size_t)descs[i].width * descs[i].height + 255) / 256 * 256;
    DevBuf d_ix(ix_total + 256);
    if (!d_ix.p) return false;
    std::vector<Q2Image> imgs; std::vector<int> which;
    uint32_t total_tiles = 0, most = 0;
    size_t ix_off = 0;
    for (int i = 0; i < n; ++i) {
        const gb200_qoix_desc& d = descs[i];
        if (d.channels < 3) continue;
        out_len[i] = 0;
        Q2Image Q;
        if (!q2_setup(Q, pixels_dev[i], d.width, d.height, d.pitchBytes, d.channels, d.bitdepth, d.colorspace, d.compression,
                      d.pixelAspectRatio, d.resolutionY, out_dev[i], d_ix.as<uint8_t>() + ix_off, total_tiles)) continue;
        ix_off += ((size_t)d.width * d.height + 255) / 256 * 256;
        imgs.push_back(Q); which.push_back(i);
        most = std::max(most, Q.base.ntiles);
    }
    const int m = (int)imgs.size();
    if (!m) return true;
    DevBuf d_imgs(sizeof(Q2Image) * (size_t)m), d_tiles(sizeof(QnTile) * ((size_t)total_tiles + 1)), d_len(sizeof(int) * (size_t)m);
    PinnedBuf h_len(sizeof(int) * (size_t)m);
    if (!d_imgs.p || !d_tiles.p || !d_len.p || !h_len.p) return false;
    bool ok = cuda_ok(cudaMemcpyAsync(d_imgs.p, imgs.data(), sizeof(Q2Image) * (size_t)m, cudaMemcpyHostToDevice, st), "q2 imgs", __FILE__, __LINE__);
    for (int k0 = 0; ok && k0 < m; k0 += 65535) {                // grid.y is limited to 65535
        const int mk = std::min(65535, m - k0);
        const dim3 grid(most, (unsigned)mk);
        const Q2Image* dI = d_imgs.as<Q2Image>() + k0; QnTile* dT = d_tiles.as<QnTile>(); int* dl = d_len.as<int>() + k0;
        q2_index_kernel<<<mk, 32, 0, st>>>(dI);
        q2_tile_ne_kernel<<<grid, QN_THREADS, 0, st>>>(dI, dT);
        q2_scan_kernel<<<mk, QN_THREADS, 0, st>>>(dI, dT, 0, dl);
        q2_tile_kernel<false><<<grid, QN_THREADS, 0, st>>>(dI, dT);
        q2_scan_kernel<<<mk, QN_THREADS, 0, st>>>(dI, dT, 1, dl);
        q2_tile_kernel<true><<<grid, QN_THREADS, 0, st>>>(dI, dT);
        count_launch(6);
    }
    ok = ok && dev_read_back_async(h_len.p, d_len.p, sizeof(int) * (size_t)m, st);
    ok = cuda_ok(cudaStreamSynchronize(st), "q2 sync", __FILE__, __LINE__) && ok;      // also before d_ix returns to the pool
    ok = ok && cuda_ok(cudaGetLastError(), "q2 kernels", __FILE__, __LINE__);
    if (ok) for (int k = 0; k < m; ++k) out_len[which[k]] = h_len.as<int>()[k];
    return ok;
}

} // namespace gb
