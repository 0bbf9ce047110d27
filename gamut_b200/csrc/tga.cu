// tga.cu -- TGA decoder on the GPU (the formats either side of the hot path; BMP was the first).
//
// Reference: loadTGA (plugins/tga.d:45-105) -> TGADecoder.getImageInfo / decodeImage (codecs/tga.d:313-588): grey,
// grey + alpha, 15/16-bit and 24/32-bit colour, colour-mapped files with 8- or 16-bit indices and 8/15/16/24/32-bit
// palette entries, each raw or run-length coded, bottom-up or top-down. The host walks the header and prepares the
// palette in output channel order (tga.cuh: tga_plan); the device does the per-pixel work: files without packets are
// one thread per pixel over the whole batch (tga_raw_kernel); run-length files take two passes: one warp per image walks
// the packet headers only and leaves a checkpoint every 256 packets (the chain is serial by the format: a packet header
// says where the next one is), then one warp per checkpoint places the pixels of its 256 packets (tga_rle_kernel).
// The row flip (:537-551) and the B/R swap (:553-565) are folded into the store / the palette.
// The encoder (saveTGA, plugins/tga.d:123-149 -> TGAEncoder, codecs/tga.d:62-292) is at the end of the file; its kernels
// and their description are in tga_encode.cuh.
#include "../../include/gamut_b200.h"
#include "batch.h"
#include "tga.cuh"
#include "tga_encode.cuh"
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
    uint32_t raw_pixels = 0, ck_total = 0, most_segs = 0; int nraw = 0;
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
            else { const uint32_t ms = tga_max_segments((uint32_t)p.w * (uint32_t)p.h); J.ck_base = ck_total; ck_total += ms; most_segs = std::max(most_segs, ms); }
            jobs.push_back(J); which.push_back(i);
        }
    host_copy_parallel(hcopies.data(), hcopies.size());
    DevBuf d_ck(sizeof(TgaCheckpoint) * ((size_t)ck_total + 1)), d_nseg(4 * (size_t)(m - nraw + 1));
    if (!d_ck.p || !d_nseg.p) { delete B; return nullptr; }
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
        for (int k0 = nraw; k0 < m; k0 += 65535) {              // grid.y is limited to 65535
            const int mk = std::min(65535, m - k0);
            const TgaJob* dj = d_jobs.as<TgaJob>() + k0; uint32_t* dn = d_nseg.as<uint32_t>() + (k0 - nraw);
            tga_rle_index_kernel<<<mk, 32, 0, st>>>(dj, d_ck.as<TgaCheckpoint>(), dn);
            tga_rle_kernel<<<dim3(most_segs, (unsigned)mk), 32, 0, st>>>(dj, d_ck.as<TgaCheckpoint>(), dn);
            count_launch(2);
        }
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

// ---- encoder -----------------------------------------------------------------------------------------------------------
namespace gb {

// Encodes n device-resident images into n device buffers (each at least gb200_tga_encode_bound bytes). out_len[i] =
// file length, 0 for an image the encoder refuses.
bool tga_encode_device(int n, const uint8_t* const* pixels_dev, const gb200_tga_desc* descs, uint8_t* const* out_dev, int* out_len,
                       cudaStream_t st)
{
    if (!ensure_device()) return false;
    std::vector<TeImage> imgs; std::vector<int> which;
    uint32_t total_rows = 0; int most = 0;
    for (int i = 0; i < n; ++i) {
        out_len[i] = 0;
        TeImage T;
        if (!te_setup(T, pixels_dev[i], descs[i].type, descs[i].width, descs[i].height, descs[i].pitchBytes, out_dev[i], total_rows)) continue;
        imgs.push_back(T); which.push_back(i);
        most = std::max(most, T.h);
    }
    const int m = (int)imgs.size();
    if (!m) return true;
    DevBuf d_imgs(sizeof(TeImage) * (size_t)m), d_bytes(4 * ((size_t)total_rows + 1)), d_off(4 * ((size_t)total_rows + 1)), d_len(sizeof(int) * (size_t)m);
    PinnedBuf h_len(sizeof(int) * (size_t)m);
    if (!d_imgs.p || !d_bytes.p || !d_off.p || !d_len.p || !h_len.p) return false;
    bool ok = cuda_ok(cudaMemcpyAsync(d_imgs.p, imgs.data(), sizeof(TeImage) * (size_t)m, cudaMemcpyHostToDevice, st), "te imgs", __FILE__, __LINE__);
    for (int k0 = 0; ok && k0 < m; k0 += 65535) {               // grid.y is limited to 65535
        const int mk = std::min(65535, m - k0);
        const dim3 grid((unsigned)most, (unsigned)mk);
        const TeImage* dI = d_imgs.as<TeImage>() + k0;
        te_row_kernel<false><<<grid, 32, 0, st>>>(dI, d_bytes.as<uint32_t>(), d_off.as<uint32_t>());
        te_scan_kernel<<<mk, 256, 0, st>>>(dI, d_bytes.as<uint32_t>(), d_off.as<uint32_t>(), d_len.as<int>() + k0);
        te_row_kernel<true><<<grid, 32, 0, st>>>(dI, d_bytes.as<uint32_t>(), d_off.as<uint32_t>());
        count_launch(3);
    }
    ok = ok && dev_read_back_async(h_len.p, d_len.p, sizeof(int) * (size_t)m, st);
    ok = cuda_ok(cudaStreamSynchronize(st), "te sync", __FILE__, __LINE__) && ok;
    ok = ok && cuda_ok(cudaGetLastError(), "te kernels", __FILE__, __LINE__);
    if (ok) for (int k = 0; k < m; ++k) out_len[which[k]] = h_len.as<int>()[k];
    return ok;
}

} // namespace gb

GB_API size_t gb200_tga_encode_bound(const gb200_tga_desc* desc)
{
    return desc ? te_bound(desc->type, desc->width, desc->height) : 0;
}

GB_API int gb200_tga_encode_batch_device(int n, const uint8_t* const* pixels_dev, const gb200_tga_desc* descs, uint8_t* const* out_dev,
                                         int* out_len, void* stream)
{
    gb::clear_error();
    if (n < 0 || !pixels_dev || !descs || !out_dev || !out_len) { gb::set_error("tga_encode_batch_device: bad arguments"); return 0; }
    return gb::tga_encode_device(n, pixels_dev, descs, out_dev, out_len, (cudaStream_t)stream) ? 1 : 0;
}

// saveTGA (plugins/tga.d:123-149): `pixels` = the first scanline of an l8 / la8 / rgb8 / rgba8 image on the host,
// desc->pitchBytes signed. malloc()'d file out (free with gb200_free), *out_len its length; NULL where saveTGA fails
// (other pixel types, a side above 65535).
GB_API uint8_t* gb200_tga_encode(const uint8_t* pixels, const gb200_tga_desc* desc, int* out_len)
{
    gb::clear_error();
    if (!gb::ensure_device()) return nullptr;
    const int sc = desc ? te_src_channels(desc->type) : 0;
    if (!pixels || !out_len || !sc || desc->width < 0 || desc->height < 0 || desc->width > 65535 || desc->height > 65535) {
        gb::set_error("tga_encode: unsupported image (TGA takes l8 / la8 / rgb8 / rgba8 up to 65535 x 65535)");
        return nullptr;
    }
    if (desc->width == 0 || desc->height == 0) {                 // the reference writes the header and no scanline bytes (:147-148)
        uint8_t* out = (uint8_t*)calloc(18, 1);
        if (!out) return nullptr;
        out[2] = 10; out[12] = (uint8_t)(desc->width & 0xff); out[13] = (uint8_t)(desc->width >> 8);
        out[14] = (uint8_t)(desc->height & 0xff); out[15] = (uint8_t)(desc->height >> 8); out[16] = (uint8_t)(((sc & 1) ? 3 : 4) * 8);
        *out_len = 18;
        return out;
    }
    const size_t row = (size_t)desc->width * sc;
    const size_t ap = desc->pitchBytes < 0 ? (size_t)(-(long long)desc->pitchBytes) : (size_t)desc->pitchBytes;
    if (ap < row && desc->height > 1) { gb::set_error("tga_encode: pitch smaller than a scanline"); return nullptr; }
    cudaStream_t st = gb::thread_stream();
    const size_t span = ap * (size_t)(desc->height - 1) + row, cap = gb200_tga_encode_bound(desc);
    if (cap > 0x7fffffffull) { gb::set_error("tga_encode: file would exceed 2 GiB"); return nullptr; }
    const uint8_t* lowest = desc->pitchBytes < 0 ? pixels - ap * (size_t)(desc->height - 1) : pixels;
    gb::DevBuf d_in(span), d_out(cap);
    if (!d_in.p || !d_out.p) return nullptr;
    if (!gb::cuda_ok(cudaMemcpyAsync(d_in.p, lowest, span, cudaMemcpyHostToDevice, st), "te h2d", __FILE__, __LINE__)) { cudaStreamSynchronize(st); return nullptr; }
    const uint8_t* pin[1] = {d_in.as<uint8_t>() + (pixels - lowest)}; uint8_t* pout[1] = {d_out.as<uint8_t>()};
    int len = 0;
    if (!gb::tga_encode_device(1, pin, desc, pout, &len, st) || len <= 0) { cudaStreamSynchronize(st); return nullptr; }
    uint8_t* out = (uint8_t*)malloc((size_t)len);
    if (!out) return nullptr;
    const bool ok = gb::cuda_ok(cudaMemcpyAsync(out, d_out.p, (size_t)len, cudaMemcpyDeviceToHost, st), "te d2h", __FILE__, __LINE__) &&
                    gb::cuda_ok(cudaStreamSynchronize(st), "te sync", __FILE__, __LINE__);
    if (!ok) { cudaStreamSynchronize(st); free(out); return nullptr; }
    *out_len = len;
    return out;
}
