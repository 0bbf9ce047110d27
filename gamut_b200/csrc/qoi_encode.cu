// qoi_encode.cu -- plain QOI encoder on the GPU (the save side of configs[0], saveQOI, plugins/qoi.d:150).
//
// Reference: qoi_encode (codecs/qoi.d:295-426). The decoder of this format is one serial chain (the index is hashed by
// pixel VALUE, so a decoder does not know which slot a pixel lands in before it has the pixel). The encoder has every
// pixel in front of it, and nothing in it is serial except where a code lands in the byte stream:
//   * a pixel equal to its predecessor joins a run; runs are cut every 62 pixels from the start of the maximal sequence
//     of such pixels and emit ONE byte at their last pixel. The position of a pixel inside its sequence is its distance
//     to the last pixel that differs from its predecessor: a prefix maximum;
//   * index[h] holds the latest pixel with hash h that was NOT a run pixel (run pixels do not touch the index, every
//     other pixel either finds itself there or is stored there, :369-377). "QOI_OP_INDEX" for pixel i therefore asks:
//     does the latest earlier non-run pixel with i's hash have i's value? With no such pixel the slot still holds its
//     initial zero (:322), which only the all-zero pixel (hash 0) can equal. That is a "last writer" query per bucket:
//     inside a tile it is answered with one 1024-bit occupancy mask per bucket in shared memory, across tiles with
//     the per-bucket last writer of every tile carried forward (64 values per tile);
//   * DIFF / LUMA / RGB / RGBA depend on pixels i and i-1 only;
//   * the byte position of a code is the prefix sum of the code lengths before it.
// Five kernels over tiles of 1024 pixels, the same shape as qoix_encode.cu: tile state (last differing pixel, last
// writer per bucket), prefix over the tiles of an image, bytes per tile, prefix sum (+ header, padding, length), emit.
// Output is byte-identical to the reference encoder (tests/test_qoi_encode_gpu.py compares with the oracle's
// restatement, itself pinned to PIL's independent QOI writer in tests/test_oracle_qoix.py; the numpy model of this file
// is tests/test_parallel_decode_models.py::_qoi_encode_model).
#include "../../include/gamut_b200.h"
#include "common.h"
#include "qoi_encode.cuh"
#include <algorithm>
#include <vector>
#include <cstring>

namespace gb {

static bool qn_valid(const gb200_qoi_desc& d) { return ::qn_valid(d.width, d.height, d.channels, d.colorspace); }

// Encodes n device-resident images into n device buffers (each at least gb200_qoi_encode_bound bytes, 16-byte
// aligned). pixels_dev[i] = first scanline, pitches[i] signed. out_len[i] = stream length, 0 for a refused image.
bool qoi_encode_device(int n, const uint8_t* const* pixels_dev, const gb200_qoi_desc* descs, const int* pitches,
                       uint8_t* const* out_dev, int* out_len, cudaStream_t st)
{
    if (!ensure_device()) return false;
    std::vector<QnImage> imgs; std::vector<int> which;
    uint32_t total_tiles = 0;
    for (int i = 0; i < n; ++i) {
        out_len[i] = 0;
        const gb200_qoi_desc& d = descs[i];
        QnImage Q;
        if (!qn_setup(Q, pixels_dev[i], d.width, d.height, d.channels, d.colorspace, pitches[i], out_dev[i], total_tiles)) continue;
        imgs.push_back(Q); which.push_back(i);
    }
    const int m = (int)imgs.size();
    if (!m) return true;
    DevBuf d_imgs(sizeof(QnImage) * (size_t)m), d_tiles(sizeof(QnTile) * ((size_t)total_tiles + 1)),
           d_val(sizeof(uint32_t) * 64 * ((size_t)total_tiles + 1)), d_len(sizeof(int) * (size_t)m);
    PinnedBuf h_len(sizeof(int) * (size_t)m);
    if (!d_imgs.p || !d_tiles.p || !d_val.p || !d_len.p || !h_len.p) return false;
    bool ok = cuda_ok(cudaMemcpyAsync(d_imgs.p, imgs.data(), sizeof(QnImage) * (size_t)m, cudaMemcpyHostToDevice, st), "qn imgs", __FILE__, __LINE__);
    if (ok) {
        uint32_t most = 0;
        for (const QnImage& Q : imgs) most = std::max(most, Q.ntiles);
        for (int k0 = 0; ok && k0 < m; k0 += 65535) {            // grid.y is limited to 65535
            const int mk = std::min(65535, m - k0);
            const dim3 grid(most, (unsigned)mk);
            const QnImage* dI = d_imgs.as<QnImage>() + k0; QnTile* dT = d_tiles.as<QnTile>(); uint32_t* dV = d_val.as<uint32_t>();
            int* dl = d_len.as<int>() + k0;
            qn_tile_state_kernel<<<grid, QN_THREADS, 0, st>>>(dI, dT, dV);
            qn_scan_kernel<<<mk, QN_THREADS, 0, st>>>(dI, dT, dV, 0, dl);
            qn_tile_kernel<false><<<grid, QN_THREADS, 0, st>>>(dI, dT, dV);
            qn_scan_kernel<<<mk, QN_THREADS, 0, st>>>(dI, dT, dV, 1, dl);
            qn_tile_kernel<true><<<grid, QN_THREADS, 0, st>>>(dI, dT, dV);
            count_launch(5);
        }
        ok = ok && dev_read_back_async(h_len.p, d_len.p, sizeof(int) * (size_t)m, st);
    }
    ok = cuda_ok(cudaStreamSynchronize(st), "qn sync", __FILE__, __LINE__) && ok;
    ok = ok && cuda_ok(cudaGetLastError(), "qn kernels", __FILE__, __LINE__);
    if (ok) for (int k = 0; k < m; ++k) out_len[which[k]] = ((const int*)h_len.p)[k];
    return ok;
}

}  // namespace gb

// qoi_encode's own allocation (:313-315) rounded up to the 16 bytes the emit kernel's word stores may touch
GB_API size_t gb200_qoi_encode_bound(const gb200_qoi_desc* desc)
{
    if (!desc) return 0;
    const unsigned long long np = (unsigned long long)desc->width * desc->height;
    return (size_t)(np * ((unsigned)desc->channels + 1u) + QOI_HEADER_SIZE + QOI_PADDING + 16);
}

GB_API int gb200_qoi_encode_batch_device(int n, const uint8_t* const* pixels_dev, const gb200_qoi_desc* descs, const int* pitches,
                                         uint8_t* const* out_dev, int* out_len, void* stream)
{
    gb::clear_error();
    if (n < 0 || !pixels_dev || !descs || !pitches || !out_dev || !out_len) { gb::set_error("qoi_encode_batch_device: bad arguments"); return 0; }
    return gb::qoi_encode_device(n, pixels_dev, descs, pitches, out_dev, out_len, (cudaStream_t)stream) ? 1 : 0;
}

// qoi_encode (qoi.d:295) as saveQOI calls it (plugins/qoi.d:150-185): `pixels` = the first scanline of an rgb8 / rgba8
// image on the host, pitchBytes signed (a vertically flipped Image has a negative pitch). malloc()'d stream out (free
// with gb200_free), *out_len its length; NULL where the reference returns null.
GB_API uint8_t* gb200_qoi_encode(const uint8_t* pixels, const gb200_qoi_desc* desc, int pitchBytes, int* out_len)
{
    gb::clear_error();
    if (!gb::ensure_device()) return nullptr;
    if (!pixels || !desc || !out_len || !gb::qn_valid(*desc)) { gb::set_error("qoi_encode: invalid image"); return nullptr; }
    const size_t row = (size_t)desc->width * desc->channels;
    const size_t ap = pitchBytes < 0 ? (size_t)(-(long long)pitchBytes) : (size_t)pitchBytes;
    if (ap < row && desc->height > 1) { gb::set_error("qoi_encode: pitch smaller than a scanline"); return nullptr; }
    cudaStream_t st = gb::thread_stream();
    // the bytes between the lowest and the highest scanline, copied as one block
    const size_t span = ap * (size_t)(desc->height - 1) + row;
    const uint8_t* lowest = pitchBytes < 0 ? pixels - ap * (size_t)(desc->height - 1) : pixels;
    const size_t cap = gb200_qoi_encode_bound(desc);
    gb::DevBuf d_in(span), d_out(cap);
    if (!d_in.p || !d_out.p) return nullptr;
    if (!gb::cuda_ok(cudaMemcpyAsync(d_in.p, lowest, span, cudaMemcpyHostToDevice, st), "qn h2d", __FILE__, __LINE__)) { cudaStreamSynchronize(st); return nullptr; }
    const uint8_t* pin[1] = {d_in.as<uint8_t>() + (pixels - lowest)}; uint8_t* pout[1] = {d_out.as<uint8_t>()};
    int len = 0;
    if (!gb::qoi_encode_device(1, pin, desc, &pitchBytes, pout, &len, st) || len <= 0) { cudaStreamSynchronize(st); return nullptr; }
    uint8_t* out = (uint8_t*)malloc((size_t)len);
    if (!out) return nullptr;
    const bool ok = gb::cuda_ok(cudaMemcpyAsync(out, d_out.p, (size_t)len, cudaMemcpyDeviceToHost, st), "qn d2h", __FILE__, __LINE__) &&
                    gb::cuda_ok(cudaStreamSynchronize(st), "qn sync", __FILE__, __LINE__);
    if (!ok) { cudaStreamSynchronize(st); free(out); return nullptr; }
    *out_len = len;
    return out;
}
