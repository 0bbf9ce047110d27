// qoix_encode.cu -- QOI-Plane10 and QOI-Plane encoders on the GPU (the other side of the config-5 path).
//
// Reference: qoiplane10_encode (codecs/qoiplane10.d:99-314) as qoix_lz4_encode calls it for 10-bit 1/2-channel images
// (plugins/qoix.d:251-339). The encoder sees every pixel at once, so -- unlike the decoder -- nothing in it is serial
// except where a code lands in the bit stream:
//   * the predictor of a pixel (locoPredict of the ORIGINAL neighbours, :84-96), its residual and its alpha
//     difference depend on four input pixels only;
//   * a pixel equal to its predecessor joins a run; runs are cut every 256 pixels from the start of the maximal
//     sequence of such pixels, and a run emits ONE code at its last pixel (a run of one whose residual is tiny
//     becomes a DIFF1). The position of a pixel inside its sequence is its distance to the last pixel that differs
//     from its predecessor: a prefix maximum;
//   * the bit position of a code is the prefix sum of the code lengths before it.
// Five kernels over tiles of 1024 pixels: last differing pixel per tile, prefix maximum over the tiles of an image,
// bits per tile, prefix sum over the tiles (+ header, end marker, length), emit. Output is bit-identical to the
// reference encoder (tests/test_qoix_encode_gpu.py compares with the oracle's restatement of it and decodes the
// result with both decoders). The LZ4 stage of qoix_lz4_encode (LZ4_compress, lz4.d:329) is not built: the stream
// is returned with compression = 0, which is what the reference itself returns whenever LZ4 does not pay.
// qoiplane_encode (codecs/qoiplane.d:109-375, 8-bit L / LA) has the same shape -- predictor = rounded-up average of the
// pixel above and the previous pixel, nibble-aligned codes, runs of at most 258 -- and runs through the same kernels
// (template parameter P8 in qoix_encode.cuh).
#include "../../include/gamut_b200.h"
#include "common.h"
#include "qoix_encode.cuh"
#include <algorithm>
#include <vector>
#include <cstring>

// qoix_encode's own checks (qoi2avg.d:386-398); the kernels are in qoi2avg_encode.cu
static bool q2_valid_desc(uint32_t width, uint32_t height, int channels, int bitdepth, int colorspace, int compression)
{
    return width && height && channels >= 3 && channels <= 4 && colorspace >= 0 && colorspace <= 2 && bitdepth == 8 && compression == 0 &&
           height < 400000000u / width;
}

namespace gb {

static bool qe_valid(const gb200_qoix_desc& d) { return ::qe_valid(d.width, d.height, d.channels, d.bitdepth, d.compression); }

template <bool P8>
static void qe_launch(const QeImage* dI, int m, uint32_t most, QeTile* dT, int* dl, cudaStream_t st)
{
    for (int k0 = 0; k0 < m; k0 += 65535) {                     // grid.y is limited to 65535
        const int mk = std::min(65535, m - k0);
        const dim3 grid(most, (unsigned)mk);
        qe_tile_ne_kernel<P8><<<grid, QE_THREADS, 0, st>>>(dI + k0, mk, dT);
        qe_scan_kernel<<<mk, QE_THREADS, 0, st>>>(dI + k0, dT, 0, dl + k0);
        qe_tile_kernel<false, P8><<<grid, QE_THREADS, 0, st>>>(dI + k0, mk, dT);
        qe_scan_kernel<<<mk, QE_THREADS, 0, st>>>(dI + k0, dT, 1, dl + k0);
        qe_tile_kernel<true, P8><<<grid, QE_THREADS, 0, st>>>(dI + k0, mk, dT);
        count_launch(5);
    }
}

// Encodes n device-resident images into n device buffers (each at least gb200_qoix_encode_bound bytes, 16-byte
// aligned). out_len[i] = stream length, 0 for an image the encoder refuses. 10-bit images go through the QOI-Plane10
// kernels, 8-bit images through the QOI-Plane ones (the table holds the 10-bit images first).
bool qoiplane_encode_device(int n, const uint8_t* const* pixels_dev, const gb200_qoix_desc* descs, uint8_t* const* out_dev,
                            int* out_len, cudaStream_t st)
{
    if (!ensure_device()) return false;
    std::vector<QeImage> imgs; std::vector<int> which;
    uint32_t total_tiles = 0, most[2] = {0, 0};
    int count[2] = {0, 0};
    for (int i = 0; i < n; ++i) out_len[i] = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (int i = 0; i < n; ++i) {
            const gb200_qoix_desc& d = descs[i];
            if ((d.bitdepth == 8) != (pass == 1)) continue;
            QeImage Q;
            if (!qe_setup(Q, pixels_dev[i], d.width, d.height, d.pitchBytes, d.channels, d.bitdepth, d.colorspace, d.compression,
                          d.pixelAspectRatio, d.resolutionY, out_dev[i], total_tiles)) continue;
            imgs.push_back(Q); which.push_back(i);
            ++count[pass]; most[pass] = std::max(most[pass], Q.ntiles);
        }
    const int m = (int)imgs.size();
    if (!m) return true;
    DevBuf d_imgs(sizeof(QeImage) * (size_t)m), d_tiles(sizeof(QeTile) * ((size_t)total_tiles + 1)), d_len(sizeof(int) * (size_t)m);
    PinnedBuf h_len(sizeof(int) * (size_t)m);
    if (!d_imgs.p || !d_tiles.p || !d_len.p || !h_len.p) return false;
    bool ok = cuda_ok(cudaMemcpyAsync(d_imgs.p, imgs.data(), sizeof(QeImage) * (size_t)m, cudaMemcpyHostToDevice, st), "qe imgs", __FILE__, __LINE__);
    if (ok) {
        if (count[0]) qe_launch<false>(d_imgs.as<QeImage>(), count[0], most[0], d_tiles.as<QeTile>(), d_len.as<int>(), st);
        if (count[1]) qe_launch<true>(d_imgs.as<QeImage>() + count[0], count[1], most[1], d_tiles.as<QeTile>(), d_len.as<int>() + count[0], st);
        ok = ok && dev_read_back_async(h_len.p, d_len.p, sizeof(int) * (size_t)m, st);
    }
    ok = cuda_ok(cudaStreamSynchronize(st), "qe sync", __FILE__, __LINE__) && ok;
    ok = ok && cuda_ok(cudaGetLastError(), "qe kernels", __FILE__, __LINE__);
    if (ok) for (int k = 0; k < m; ++k) out_len[which[k]] = ((const int*)h_len.p)[k];
    return ok;
}

// qoi2avg_encode.cu: the 3 / 4-channel 8-bit images of a batch (QOI2AVG)
bool qoi2avg_encode_device(int n, const uint8_t* const* pixels_dev, const gb200_qoix_desc* descs, uint8_t* const* out_dev,
                           int* out_len, cudaStream_t st);
// qoi10b_encode.cu: the 3 / 4-channel 10-bit images of a batch (QOI-10b)
bool qoi10b_encode_device(int n, const uint8_t* const* pixels_dev, const gb200_qoix_desc* descs, uint8_t* const* out_dev,
                          int* out_len, cudaStream_t st);

}  // namespace gb

// worst case of qoiplane10_encode's own allocation (:112-116; qoiplane_encode's, qoiplane.d:125-129, is smaller) rounded
// up for the word stores of the emit kernel
GB_API size_t gb200_qoix_encode_bound(const gb200_qoix_desc* desc)
{
    if (!desc) return 0;
    const unsigned long long np = (unsigned long long)desc->width * desc->height;
    if (desc->channels >= 3 && desc->bitdepth == 10) return (size_t)((np * 52 + 7) / 8 + QOIX_HEADER_SIZE + 5 + 64); // QOI-10b: at most 52 bits per pixel
    if (desc->channels >= 3) return (size_t)(np * ((unsigned)desc->channels + 1u) + QOIX_HEADER_SIZE + 4 + 16);     // qoix_encode's own (qoi2avg.d:408)
    return (size_t)((np * (desc->channels == 1 ? 14 : 28) + 7) / 8 + QOIX_HEADER_SIZE + 5 + 64);
}

GB_API int gb200_qoix_encode_batch_device(int n, const uint8_t* const* pixels_dev, const gb200_qoix_desc* descs,
                                          uint8_t* const* out_dev, int* out_len, void* stream)
{
    gb::clear_error();
    if (n < 0 || !pixels_dev || !descs || !out_dev || !out_len) { gb::set_error("qoix_encode_batch_device: bad arguments"); return 0; }
    bool any_rgb8 = false, any_rgb10 = false;
    for (int i = 0; i < n; ++i) if (descs[i].channels >= 3) { if (descs[i].bitdepth == 10) any_rgb10 = true; else any_rgb8 = true; }
    if (!gb::qoiplane_encode_device(n, pixels_dev, descs, out_dev, out_len, (cudaStream_t)stream)) return 0;       // 1 / 2 channels; others get 0
    if (any_rgb8 && !gb::qoi2avg_encode_device(n, pixels_dev, descs, out_dev, out_len, (cudaStream_t)stream)) return 0;
    if (any_rgb10 && !gb::qoi10b_encode_device(n, pixels_dev, descs, out_dev, out_len, (cudaStream_t)stream)) return 0;
    return 1;
}

// qoix_lz4_encode (plugins/qoix.d:251-339) with its dispatch (:268-290): 10-bit with 1 / 2 channels -> qoiplane10_encode,
// 10-bit with 3 / 4 -> qoi10b_encode, 8-bit with 1 / 2 -> qoiplane_encode, 8-bit with 3 / 4 -> qoix_encode (QOI2AVG). Host
// pixels in, malloc()'d stream out (free with gb200_free), *out_len its length. The stream is not LZ4-wrapped
// (compression 0).
GB_API uint8_t* gb200_qoix_encode(const uint8_t* pixels, const gb200_qoix_desc* desc, int* out_len)
{
    gb::clear_error();
    if (!gb::ensure_device()) return nullptr;
    const bool rgb8 = desc && desc->channels >= 3 && ::q2_valid_desc(desc->width, desc->height, desc->channels, desc->bitdepth, desc->colorspace, desc->compression);
    const bool rgb10 = desc && desc->channels >= 3 && desc->channels <= 4 && desc->bitdepth == 10 && desc->compression == 0 && desc->width && desc->height &&
                       desc->height < 400000000u / desc->width && (unsigned long long)desc->width * desc->height * 52ull + 4096 < 0xffffffffull;   // qoi10b.d:138-146
    if (!pixels || !desc || !out_len || !(rgb8 || rgb10 || gb::qe_valid(*desc)) ||
        desc->pitchBytes < (int)(desc->width * desc->channels * (desc->bitdepth == 10 ? 2u : 1u))) {
        gb::set_error("qoix_encode: unsupported image (built: QOI-Plane10 / QOI-Plane for 10-bit / 8-bit images with 1 or 2 channels, QOI-10b / QOI2AVG for 10-bit / 8-bit images with 3 or 4)");
        return nullptr;
    }
    cudaStream_t st = gb::thread_stream();
    const size_t in_bytes = (size_t)desc->pitchBytes * desc->height, cap = gb200_qoix_encode_bound(desc);
    gb::DevBuf d_in(in_bytes), d_out(cap);
    if (!d_in.p || !d_out.p) return nullptr;
    if (!gb::cuda_ok(cudaMemcpyAsync(d_in.p, pixels, in_bytes, cudaMemcpyHostToDevice, st), "qe h2d", __FILE__, __LINE__)) { cudaStreamSynchronize(st); return nullptr; }
    const uint8_t* pin[1] = {d_in.as<uint8_t>()}; uint8_t* pout[1] = {d_out.as<uint8_t>()};
    int len = 0;
    const bool ran = rgb8 ? gb::qoi2avg_encode_device(1, pin, desc, pout, &len, st) : rgb10 ? gb::qoi10b_encode_device(1, pin, desc, pout, &len, st)
                          : gb::qoiplane_encode_device(1, pin, desc, pout, &len, st);
    if (!ran || len <= 0) { cudaStreamSynchronize(st); return nullptr; }
    uint8_t* out = (uint8_t*)malloc((size_t)len);
    if (!out) return nullptr;
    const bool ok = gb::cuda_ok(cudaMemcpyAsync(out, d_out.p, (size_t)len, cudaMemcpyDeviceToHost, st), "qe d2h", __FILE__, __LINE__) &&
                    gb::cuda_ok(cudaStreamSynchronize(st), "qe sync", __FILE__, __LINE__);
    if (!ok) { cudaStreamSynchronize(st); free(out); return nullptr; }
    *out_len = len;
    return out;
}
