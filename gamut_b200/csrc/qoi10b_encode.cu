// qoi10b_encode.cu -- QOI-10b encoder on the GPU: qoi10b_encode (codecs/qoi10b.d:136-500), the codec
// qoix_lz4_encode / saveQOIX pick for rgb16 / rgba16 / rgbap16 images (plugins/qoix.d:213-228, :268-278). Kernels and
// their description: qoi10b_encode.cuh. Reached through gb200_qoix_encode / gb200_qoix_encode_batch_device
// (qoix_encode.cu) for descs with bitdepth 10 and 3 or 4 channels. The LZ4 stage is not built: compression = 0.
#include "../../include/gamut_b200.h"
#include "common.h"
#include "qoi10b_encode.cuh"
#include <algorithm>
#include <vector>
#include <cstring>

namespace gb {

// Encodes the 3 / 4-channel 10-bit images of a batch (other entries are left alone: out_len[i] is only written for
// those). out_dev[i]: at least gb200_qoix_encode_bound bytes, 16-byte aligned.
bool qoi10b_encode_device(int n, const uint8_t* const* pixels_dev, const gb200_qoix_desc* descs, uint8_t* const* out_dev,
                          int* out_len, cudaStream_t st)
{
    if (!ensure_device()) return false;
    std::vector<QeImage> imgs; std::vector<int> which;
    uint32_t total_tiles = 0, most = 0;
    for (int i = 0; i < n; ++i) {
        const gb200_qoix_desc& d = descs[i];
        if (d.channels < 3 || d.bitdepth != 10) continue;
        out_len[i] = 0;
        QeImage Q;
        if (!q10_setup(Q, pixels_dev[i], d.width, d.height, d.pitchBytes, d.channels, d.bitdepth, d.colorspace, d.compression,
                       d.pixelAspectRatio, d.resolutionY, out_dev[i], total_tiles)) continue;
        imgs.push_back(Q); which.push_back(i);
        most = std::max(most, Q.ntiles);
    }
    const int m = (int)imgs.size();
    if (!m) return true;
    DevBuf d_imgs(sizeof(QeImage) * (size_t)m), d_tiles(sizeof(QeTile) * ((size_t)total_tiles + 1)), d_len(sizeof(int) * (size_t)m);
    PinnedBuf h_len(sizeof(int) * (size_t)m);
    if (!d_imgs.p || !d_tiles.p || !d_len.p || !h_len.p) return false;
    bool ok = cuda_ok(cudaMemcpyAsync(d_imgs.p, imgs.data(), sizeof(QeImage) * (size_t)m, cudaMemcpyHostToDevice, st), "q10 imgs", __FILE__, __LINE__);
    for (int k0 = 0; ok && k0 < m; k0 += 65535) {                // grid.y is limited to 65535
        const int mk = std::min(65535, m - k0);
        const dim3 grid(most, (unsigned)mk);
        const QeImage* dI = d_imgs.as<QeImage>() + k0; QeTile* dT = d_tiles.as<QeTile>(); int* dl = d_len.as<int>() + k0;
        q10_tile_ne_kernel<<<grid, QE_THREADS, 0, st>>>(dI, dT);
        qe_scan_kernel<<<mk, QE_THREADS, 0, st>>>(dI, dT, 0, dl);
        q10_tile_kernel<false><<<grid, QE_THREADS, 0, st>>>(dI, dT);
        qe_scan_kernel<<<mk, QE_THREADS, 0, st>>>(dI, dT, 1, dl);
        q10_tile_kernel<true><<<grid, QE_THREADS, 0, st>>>(dI, dT);
        count_launch(5);
    }
    ok = ok && dev_read_back_async(h_len.p, d_len.p, sizeof(int) * (size_t)m, st);
    ok = cuda_ok(cudaStreamSynchronize(st), "q10 sync", __FILE__, __LINE__) && ok;
    ok = ok && cuda_ok(cudaGetLastError(), "q10 kernels", __FILE__, __LINE__);
    if (ok) for (int k = 0; k < m; ++k) out_len[which[k]] = h_len.as<int>()[k];
    return ok;
}

} // namespace gb
