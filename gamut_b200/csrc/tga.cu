// tga.cu -- TGA decoder on the GPU (the formats either side of the hot path; BMP was the first).
//
// Reference: loadTGA (plugins/tga.d:45-105) -> TGADecoder.getImageInfo / decodeImage (codecs/tga.d:313-588): grey,
// grey + alpha, 15/16-bit and 24/32-bit colour, colour-mapped files with 8- or 16-bit indices and 8/15/16/24/32-bit
// palette entries, each raw or run-length coded, bottom-up or top-down. The host walks the header and prepares the
// palette in output channel order (tga.cuh: tga_plan); the device does the per-pixel work: files without packets are
// one thread per pixel over the whole batch (tga_raw_kernel); run-length files are one warp per image, the packet
// chain walked by the warp and the pixels of a packet placed by its lanes (tga_rle_kernel) -- the chain is serial by
// the format (a packet header says where the next one is), so this half is a parity path, not a fast one.
// The row flip (:537-551) and the B/R swap (:553-565) are folded into the store / the palette.
#include "../../include/gamut_b200.h"
#include "batch.h"
#include "tga.cuh"
#include <algorithm>
#include <chrono>
#include <cstring>

namespace gb {

namespace {
inline size_t al(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
inline double now_ms() { using namespace std::chrono; return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count(); }
}

// Decodes n files. files[i] = host bytes (always needed: the header and the palette are read on the host); files_dev,
// when given, holds the same bytes on the device. Images that fail have status 0.
gb200_batch* tga_decode_batch(int n, const uint8_t* const* files, const size_t* lens, const uint8_t* const* files_dev, cudaStream_t st)
{
    if (!ensure_device()) return nullptr;
    if (n < 0 || (n > 0 && (!files || !lens))) { set_error("tga_decode_batch: bad arguments"); return nullptr; }
    gb200_batch* B = new gb200_batch;
    B->stream = st;
    B->images.resize((size_t)n);
    for (auto& D : B->images) { memset(&D, 0, sizeof(D)); D.ppmX = D.ppmY = D.pixelAspectRatio = -1; }
    const double t0 = now_ms();
    std::vector<TgaPlan> P((size_t)n);
    std::vector<int> live;
    size_t out_total = 0, stage_total = 0;
    std::vector<size_t> out_off((size_t)n, 0), file_off((size_t)n, 0), pal_off((size_t)n, 0);
    uint64_t total_pixels = 0;
    for (int i = 0; i < n; ++i) {
        if (!files[i] || lens[i] > 0xfffffff0u || !tga_plan(files[i], lens[i], P[i])) continue;
        const uint64_t px = (uint64_t)P[i].w * P[i].h;
        if (total_pixels + px > 0xfffffff0ull) { P[i].ok = false; continue; }
        live.push_back(i);
        total_pixels += px;
        out_off[i] = out_total; out_total += al((size_t)px * P[i].components);
        if (!files_dev) { file_off[i] = stage_total; stage_total += al(lens[i] + 16); }
        if (!P[i].palette.empty()) { pal_off[i] = stage_total; stage_total += al(P[i].palette.size()); }
    }
    B->host_parse_ms = now_ms() - t0;
    const int m = (int)live.size();
    if (!m) return B;
    uint8_t* d_out = (uint8_t*)dev_alloc(out_total);
    if (!d_out) { delete B; return nullptr; }
    B->device_allocs.push_back(d_out);
    DevBuf d_stage(stage_total + 256), d_jobs(sizeof(TgaJob) * (size_t)m), d_fail(sizeof(int) * (size_t)m);
    PinnedBuf h_stage(stage_total + 256), h_fail(sizeof(int) * (size_t)m);
    if (!d_stage.p || !d_jobs.p || !d_fail.p || !h_stage.p || !h_fail.p) { delete B; return nullptr; }
    // the table holds the files without packets first (one flat launch over their pixels), then the run-length files
    std::vector<TgaJob> jobs; std::vector<int> which;
    std::vector<HostCopy> hcopies;
    uint32_t raw_pixels = 0; int nraw = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (int i : live) {
            const TgaPlan& p = P[i];
            if ((p.rle != 0) != (pass == 1)) continue;
            TgaJob J; memset(&J, 0, sizeof(J));
            if (files_dev) J.data = files_dev[i];
            else { hcopies.push_back(HostCopy{h_stage.as<uint8_t>() + file_off[i], files[i], lens[i]}); J.data = d_stage.as<uint8_t>() + file_off[i]; }
            if (!p.palette.empty()) {
                hcopies.push_back(HostCopy{h_stage.as<uint8_t>() + pal_off[i], p.palette.data(), p.palette.size()});
                J.palette = d_stage.as<uint8_t>() + pal_off[i];
            }
            J.out = d_out + out_off[i];
            J.fail = d_fail.as<int>() + (int)jobs.size();
            J.len = (uint32_t)lens[i]; J.pix_off = p.pix_off; J.palette_len = p.palette_len;
            J.w = p.w; J.h = p.h; J.components = p.components; J.src_bytes = p.src_bytes; J.mode = p.mode; J.index16 = p.index16;
            J.inverted = p.inverted; J.rle = p.rle;
            if (!p.rle) { J.pix_base = raw_pixels; raw_pixels += (uint32_t)p.w * (uint32_t)p.h; ++nraw; }
            jobs.push_back(J); which.push_back(i);
        }
    host_copy_parallel(hcopies.data(), hcopies.size());
    cudaEvent_t ev[3];
    for (auto& e : ev) cudaEventCreate(&e);
    bool okc = true;
    cudaEventRecord(ev[0], st);
    if (stage_total) okc &= cuda_ok(cudaMemcpyAsync(d_stage.p, h_stage.p, stage_total, cudaMemcpyHostToDevice, st), "tga files", __FILE__, __LINE__);
    okc &= cuda_ok(cudaMemcpyAsync(d_jobs.p, jobs.data(), sizeof(TgaJob) * (size_t)m, cudaMemcpyHostToDevice, st), "tga jobs", __FILE__, __LINE__);
    okc = okc && dev_fill_async(d_fail.p, 0, sizeof(int) * (size_t)m, st);
    cudaEventRecord(ev[1], st);
    if (okc) {
        if (nraw) { tga_raw_kernel<<<(raw_pixels + 255) / 256, 256, 0, st>>>(d_jobs.as<TgaJob>(), nraw, raw_pixels); count_launch(); }
        if (m > nraw) { tga_rle_kernel<<<m - nraw, 32, 0, st>>>(d_jobs.as<TgaJob>() + nraw); count_launch(); }
        okc = dev_read_back_async(h_fail.p, d_fail.p, sizeof(int) * (size_t)m, st);
    }
    cudaEventRecord(ev[2], st);
    okc &= cuda_ok(cudaStreamSynchronize(st), "tga sync", __FILE__, __LINE__);
    okc &= cuda_ok(cudaGetLastError(), "tga kernels", __FILE__, __LINE__);
    if (okc) for (int q = 0; q < 2; ++q) { float ms = 0; cudaEventElapsedTime(&ms, ev[q], ev[q + 1]); B->phase_ms[q] += ms; }
    for (auto& e : ev) cudaEventDestroy(e);
    if (!okc) { cudaStreamSynchronize(st); delete B; return nullptr; }
    for (int k = 0; k < m; ++k) {
        if (h_fail.as<int>()[k]) continue;                       // a read past the end of the file: decodeImage returns null
        const int i = which[k];
        gb200_image_desc& D = B->images[i];
        const TgaPlan& p = P[i];
        D.status = 1; D.pixels = d_out + out_off[i];
        D.width = p.w; D.height = p.h; D.channels = p.components; D.file_channels = p.components; D.bits = 8;
        D.pixel_type = p.components == 1 ? GB200_l8 : p.components == 2 ? GB200_la8 : p.components == 3 ? GB200_rgb8 : GB200_rgba8;   // plugins/tga.d:72-79
        D.pitch = p.w * p.components;
    }
    B->device_ms = now_ms() - t0 - B->host_parse_ms;
    return B;
}

} // namespace gb

GB_API gb200_batch* gb200_tga_decode_batch(int n, const uint8_t* const* files, const size_t* lens,
                                           const uint8_t* const* files_dev, void* stream)
{
    gb::clear_error();
    return gb::tga_decode_batch(n, files, lens, files_dev, (cudaStream_t)stream);
}

// TGADecoder.getImageInfo + decodeImage (codecs/tga.d:313-588) as loadTGA calls them (plugins/tga.d:45-70): host bytes in,
// malloc'd host pixels out, *comp = components (1 = l8, 2 = la8, 3 = rgb8, 4 = rgba8).
GB_API uint8_t* gb200_tga_load(const uint8_t* data, size_t len, int* width, int* height, int* comp)
{
    gb::clear_error();
    if (!gb::ensure_device()) return nullptr;
    const uint8_t* f[1] = {data}; size_t l[1] = {len};
    cudaStream_t st = gb::thread_stream();
    gb200_batch* B = gb::tga_decode_batch(1, f, l, nullptr, st);
    if (!B) return nullptr;
    const gb200_image_desc& D = B->images[0];
    if (!D.status) { gb::set_error("TGA decoding failed"); delete B; return nullptr; }
    const size_t bytes = (size_t)D.pitch * D.height;
    uint8_t* out = (uint8_t*)malloc(bytes ? bytes : 1);
    const bool ok = out && gb::cuda_ok(cudaMemcpyAsync(out, D.pixels, bytes, cudaMemcpyDeviceToHost, st), "d2h", __FILE__, __LINE__) &&
                    gb::cuda_ok(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__);
    if (width) *width = D.width;
    if (height) *height = D.height;
    if (comp) *comp = D.channels;
    delete B;
    if (!ok) { free(out); return nullptr; }
    return out;
}
