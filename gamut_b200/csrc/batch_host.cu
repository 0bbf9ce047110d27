// batch_host.cu -- host-to-host batched decode with the transfers overlapped.
//
// gb200_decode_batch_host is what a caller with files in host memory and pixels wanted in host memory uses (the shape
// of the reference's codec calls: bytes in, malloc'd pixels out -- plugins/png.d:108, jpeg.d:62, qoix.d:116 -- for a
// whole batch). The batch is cut into sub-batches; while the kernels of sub-batch k+1 run (and its files travel to the
// device), the pixels of sub-batch k travel back on a second stream: PCIe is full duplex and the copy engines run
// beside the SMs, so the end-to-end time approaches max(download, decode) instead of their sum.
#include "common.h"
#include "batch.h"
#include <vector>
#include <memory>

namespace gb {
gb200_batch* png_decode_batch(int n, const uint8_t* const* files, const size_t* lens, const uint8_t* const* files_dev,
                              int req_comp, int want16, cudaStream_t st);
gb200_batch* jpeg_decode_batch(int n, const uint8_t* const* files, const size_t* lens, const uint8_t* const* files_dev,
                               int req_comps, cudaStream_t st);
gb200_batch* qoix_decode_batch(int n, const uint8_t* const* files, const size_t* lens, const uint8_t* const* files_dev,
                               int flags, cudaStream_t st);
}

GB_API int gb200_decode_batch_host(int format, int n, const uint8_t* const* files, const size_t* lens, int arg, int want16,
                                   uint8_t* dst_host, size_t dst_stride, gb200_image_desc* descs, int sub_batch)
{
    gb::clear_error();
    if (!gb::ensure_device()) return 0;
    if (n < 0 || (n > 0 && (!files || !lens || !dst_host || !descs))) { gb::set_error("gb200_decode_batch_host: bad arguments"); return 0; }
    if (format != GB200_FORMAT_JPEG && format != GB200_FORMAT_PNG && format != GB200_FORMAT_QOIX) {
        gb::set_error("gb200_decode_batch_host: format %d has no batched decoder", format); return 0;
    }
    if (sub_batch <= 0) {
        // JPEG decode time scales with the number of images, so many small sub-batches overlap well. The PNG and QOIX
        // pipelines have per-stream serial stages (LZ77 resolve, chain walks) whose duration hardly depends on how
        // many streams run side by side: cutting those batches only adds latency, so they stay whole up to 1024 images.
        if (format == GB200_FORMAT_JPEG) { sub_batch = n / 8; if (sub_batch < 8) sub_batch = 8; if (sub_batch > 64) sub_batch = 64; }
        else sub_batch = 1024;
    }
    cudaStream_t s_decode = gb::thread_stream(0), s_copy = gb::thread_stream(1);
    if (!s_decode || !s_copy) return 0;
    struct InFlight { gb200_batch* B = nullptr; cudaEvent_t done = nullptr; };
    InFlight fly[2];
    bool ok = true;
    auto retire = [&](InFlight& f) {
        if (!f.B) return;
        if (f.done) { if (cudaEventSynchronize(f.done) != cudaSuccess) ok = false; cudaEventDestroy(f.done); f.done = nullptr; }
        delete f.B; f.B = nullptr;
    };
    int slot = 0;
    for (int a = 0; a < n && ok; a += sub_batch, slot ^= 1) {
        const int m = n - a < sub_batch ? n - a : sub_batch;
        retire(fly[slot]);                       // the sub-batch before the previous one: its pixels have long arrived
        gb200_batch* B = nullptr;
        switch (format) {
        case GB200_FORMAT_JPEG: B = gb::jpeg_decode_batch(m, files + a, lens + a, nullptr, arg, s_decode); break;
        case GB200_FORMAT_PNG:  B = gb::png_decode_batch(m, files + a, lens + a, nullptr, arg, want16, s_decode); break;
        default:                B = gb::qoix_decode_batch(m, files + a, lens + a, nullptr, arg, s_decode); break;
        }
        if (!B) { ok = false; break; }
        // the decode call returns with its work complete (it reads the statuses back), so the copies can be queued on
        // the other stream right away; they overlap the next sub-batch's upload and kernels
        for (int i = 0; i < m; ++i) {
            gb200_image_desc D = B->images[i];
            uint8_t* dst = dst_host + (size_t)(a + i) * dst_stride;
            if (D.status && D.pixels) {
                const size_t bytes = (size_t)D.pitch * D.height;
                if (bytes > dst_stride) { D.status = 0; }
                else if (!gb::cuda_ok(cudaMemcpyAsync(dst, D.pixels, bytes, cudaMemcpyDeviceToHost, s_copy), "pixels to host", __FILE__, __LINE__)) ok = false;
            }
            D.pixels = D.status ? dst : nullptr;
            descs[a + i] = D;
        }
        fly[slot].B = B;
        if (cudaEventCreateWithFlags(&fly[slot].done, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventRecord(fly[slot].done, s_copy) != cudaSuccess) ok = false;
    }
    retire(fly[0]); retire(fly[1]);
    if (!ok && !gb200_last_error()[0]) gb::set_error("gb200_decode_batch_host: a transfer failed");
    return ok ? 1 : 0;
}
