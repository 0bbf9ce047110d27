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
#include <string>
#include <thread>
#include <memory>
#include <chrono>
#include <stdlib.h>
#include <stdio.h>

namespace gb {
gb200_batch* png_decode_batch(int n, const uint8_t* const* files, const size_t* lens, const uint8_t* const* files_dev,
                              int req_comp, int want16, cudaStream_t st);
gb200_batch* jpeg_decode_batch(int n, const uint8_t* const* files, const size_t* lens, const uint8_t* const* files_dev,
                               int req_comps, cudaStream_t st);
gb200_batch* qoix_decode_batch(int n, const uint8_t* const* files, const size_t* lens, const uint8_t* const* files_dev,
                               int flags, cudaStream_t st);
gb200_batch* bmp_decode_batch(int n, const uint8_t* const* files, const size_t* lens, const uint8_t* const* files_dev,
                              int req_comp, cudaStream_t st);
}

// one pipeline over files [0, n): decode calls on s_decode, downloads on s_copy
static int decode_batch_host_range(int format, int n, const uint8_t* const* files, const size_t* lens, int arg, int want16,
                                   uint8_t* dst_host, size_t dst_stride, gb200_image_desc* descs, int sub_batch,
                                   cudaStream_t s_decode, cudaStream_t s_copy)
{
    if (sub_batch <= 0) {
        // The download of sub-batch k overlaps the upload + kernels of sub-batch k+1, so the call costs about one
        // (small) first decode plus the larger of the two sums. Sizes measured on B200 + PCIe 5 (profiles/r2_e2e_*):
        // JPEG 4K: 32 images (6 ms of decode against 15 ms of download); PNG / QOIX: their LZ77 stages need a few
        // hundred streams to fill the machine, so the sub-batches are larger (PNG 256: 4311 -> 4674 Mpx/s on 512 1080p files; QOIX 128: 7340 -> 7630 on 256 files). GB200_E2E_SUB overrides (measurement).
        const char* e = getenv("GB200_E2E_SUB");
        const int env = e ? atoi(e) : 0;
        if (env > 0) sub_batch = env;
        else if (format == GB200_FORMAT_JPEG) { sub_batch = n / 8; if (sub_batch < 8) sub_batch = 8; if (sub_batch > 32) sub_batch = 32; }
        else if (format == GB200_FORMAT_BMP) sub_batch = 16;
        else if (format == GB200_FORMAT_PNG) sub_batch = 256;
        else sub_batch = 64;        // QOIX: 64 since its decode got faster than its download (256 files: 128 -> 113.9 ms, 64 -> 107.8, 48 -> 115.4)
    }
    struct InFlight { gb200_batch* B = nullptr; cudaEvent_t done = nullptr; cudaEvent_t start = nullptr; double t_dec0 = 0, t_dec1 = 0; int a = 0; };
    const bool trace = getenv("GB200_E2E_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_base = now();
    cudaEvent_t ev_base = nullptr;
    if (trace) { cudaEventCreate(&ev_base); cudaEventRecord(ev_base, s_copy); }
    InFlight fly[2];
    bool ok = true;
    auto retire = [&](InFlight& f) {
        if (!f.B) return;
        if (f.done) {
            if (cudaEventSynchronize(f.done) != cudaSuccess) ok = false;
            if (trace && f.start) {
                float c0 = 0, c1 = 0;
                cudaEventElapsedTime(&c0, ev_base, f.start); cudaEventElapsedTime(&c1, ev_base, f.done);
                fprintf(stderr, "[e2e] sub-batch at %d: decode call %.2f..%.2f ms (host clock), download %.2f..%.2f ms (device clock from the first record), retired at %.2f\n",
                        f.a, f.t_dec0 - t_base, f.t_dec1 - t_base, c0, c1, now() - t_base);
                cudaEventDestroy(f.start); f.start = nullptr;
            }
            cudaEventDestroy(f.done); f.done = nullptr;
        }
        delete f.B; f.B = nullptr;
    };
    int slot = 0;
    // the first sub-batch is the only one whose decode nothing hides: a quarter of the others
    int m = 0;
    for (int a = 0; a < n && ok; a += m, slot ^= 1) {
        const int want = a == 0 && n > sub_batch ? (sub_batch + 3) / 4 : sub_batch;
        m = n - a < want ? n - a : want;
        retire(fly[slot]);                       // the sub-batch before the previous one: its pixels have long arrived
        gb200_batch* B = nullptr;
        const double td0 = now();
        switch (format) {
        case GB200_FORMAT_JPEG: B = gb::jpeg_decode_batch(m, files + a, lens + a, nullptr, arg, s_decode); break;
        case GB200_FORMAT_PNG:  B = gb::png_decode_batch(m, files + a, lens + a, nullptr, arg, want16, s_decode); break;
        case GB200_FORMAT_BMP:  B = gb::bmp_decode_batch(m, files + a, lens + a, nullptr, arg, s_decode); break;
        default:                B = gb::qoix_decode_batch(m, files + a, lens + a, nullptr, arg, s_decode); break;
        }
        if (!B) { ok = false; break; }
        fly[slot].t_dec0 = td0; fly[slot].t_dec1 = now(); fly[slot].a = a;
        if (trace) { cudaEventCreate(&fly[slot].start); cudaEventRecord(fly[slot].start, s_copy); }
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
        if (cudaEventCreateWithFlags(&fly[slot].done, trace ? cudaEventDefault : cudaEventDisableTiming) != cudaSuccess ||
            cudaEventRecord(fly[slot].done, s_copy) != cudaSuccess) ok = false;
    }
    retire(fly[0]); retire(fly[1]);
    if (!ok && !gb200_last_error()[0]) gb::set_error("gb200_decode_batch_host: a transfer failed");
    return ok ? 1 : 0;
}

GB_API int gb200_decode_batch_host(int format, int n, const uint8_t* const* files, const size_t* lens, int arg, int want16,
                                   uint8_t* dst_host, size_t dst_stride, gb200_image_desc* descs, int sub_batch)
{
    gb::clear_error();
    if (!gb::ensure_device()) return 0;
    if (n < 0 || (n > 0 && (!files || !lens || !dst_host || !descs))) { gb::set_error("gb200_decode_batch_host: bad arguments"); return 0; }
    if (format != GB200_FORMAT_JPEG && format != GB200_FORMAT_PNG && format != GB200_FORMAT_QOIX && format != GB200_FORMAT_BMP) {
        gb::set_error("gb200_decode_batch_host: format %d has no batched decoder", format); return 0;
    }
    // PNG: a decode call costs about 40 ms however few streams it holds (the LZ77 resolver is a chain of dependent
    // chunks per stream) and leaves most of the machine idle while it does, so the batch is cut into slices that run
    // their pipelines side by side from as many host threads, each on its own pair of streams (512 1080p files: one
    // pipeline 207 ms, two slices 177, three 160, four 153). The other formats fill the machine with one call (QOIX
    // two slices: 110 -> 138 ms).
    int parts = 1;
    if (format == GB200_FORMAT_PNG && sub_batch <= 0 && n >= 96) parts = n >= 128 ? 4 : 3;
    if (const char* e = getenv("GB200_E2E_PARTS")) { const int v = atoi(e); if (v >= 1 && v <= 4) parts = v; }
    if (parts > n) parts = n > 0 ? n : 1;
    cudaStream_t st[8];
    for (int i = 0; i < 2 * parts; ++i) { st[i] = gb::thread_stream(i); if (!st[i]) return 0; }
    if (parts == 1) return decode_batch_host_range(format, n, files, lens, arg, want16, dst_host, dst_stride, descs, sub_batch, st[0], st[1]);
    int dev = 0;
    cudaGetDevice(&dev);
    std::vector<int> rc((size_t)parts, 0);
    std::vector<std::string> errs((size_t)parts);
    auto run = [&](int t) {
        const int a = (int)((long long)n * t / parts), b = (int)((long long)n * (t + 1) / parts);
        if (t > 0) { cudaSetDevice(dev); gb::clear_error(); }      // the current device is per host thread
        rc[t] = decode_batch_host_range(format, b - a, files + a, lens + a, arg, want16, dst_host + (size_t)a * dst_stride, dst_stride,
                                        descs + a, sub_batch, st[2 * t], st[2 * t + 1]);
        if (!rc[t]) errs[t] = gb200_last_error();
    };
    std::vector<std::thread> th;
    for (int t = 1; t < parts; ++t) th.emplace_back(run, t);
    run(0);
    for (auto& x : th) x.join();
    for (int t = 0; t < parts; ++t) if (!rc[t]) { gb::set_error("%s", errs[t].empty() ? "gb200_decode_batch_host: a slice failed" : errs[t].c_str()); return 0; }
    return 1;
}
