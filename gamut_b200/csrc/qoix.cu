// qoix.cu -- QOI, the QOIX container (+LZ4) and its sub-codecs for sm_100a.
//
// Drop-in for qoi_decode (source/gamut/codecs/qoi.d:448-550) and qoix_lz4_decode
// (source/gamut/plugins/qoix.d:350-473) with its sub-decoders qoiplane10_decode
// (codecs/qoiplane10.d:317-515), LZ4_decompress_fast (codecs/lz4.d:976 -> :760-963).
// The 14/25-byte headers are parsed on the host; all byte-stream and pixel work runs on the GPU:
//   lz4_kernel          one warp per LZ4 block: the token walk is warp-uniform, literal runs and
//                       matches are copied 32 bytes per step by the whole warp
//   qoiplane10_kernel   QOI-Plane10 opcode stream -> 10-bit L/LA expanded to 16 bit (one thread per
//                       image: the MED predictor makes every pixel depend on its left/top/top-left)
//   qoi_kernel          QOI opcode stream (value-hashed index => serial per image)
// Unlike the reference's "fast" LZ4 variant and unchecked opcode readers, every read is bounds-checked;
// a corrupt stream fails that image only.
#include "common.h"
#include "batch.h"
#include <vector>
#include <chrono>

namespace {

constexpr int QOIX_HEADER_SIZE = 25;

struct Lz4Job { const uint8_t* in; uint32_t in_len; uint8_t* out; uint32_t orig; int image; };

// LZ4 block decode with endOnOutputSize semantics (lz4.d:760-963): stops when exactly `orig` bytes
// have been produced by a final literal run.
__global__ void __launch_bounds__(128)
lz4_kernel(const Lz4Job* __restrict__ jobs, int njobs, int* status)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= njobs) return;
    const Lz4Job J = jobs[warp];
    const uint8_t* ip = J.in; const uint8_t* const iend = J.in + J.in_len;
    uint8_t* op = J.out; uint8_t* const oend = J.out + J.orig;
    bool ok = true;
    if (J.orig == 0) { ok = J.in_len >= 1 && ip[0] == 0; goto done; }
    for (;;) {
        if (ip >= iend) { ok = false; break; }
        const uint32_t token = *ip++;
        size_t length = token >> 4;
        if (length == 15) {
            uint32_t s;
            do { if (ip >= iend) { ok = false; break; } s = *ip++; length += s; } while (s == 255);
            if (!ok) break;
        }
        if (length > (size_t)(oend - op) || length > (size_t)(iend - ip)) { ok = false; break; }
        const bool last = (op + length) + 8 > oend;          // cpy > oend - COPYLENGTH
        if (last && op + length != oend) { ok = false; break; }
        for (size_t i = lane; i < length; i += 32) op[i] = ip[i];
        ip += length; op += length;
        if (last) break;
        if (iend - ip < 2) { ok = false; break; }
        const size_t offset = (size_t)ip[0] | ((size_t)ip[1] << 8); ip += 2;
        if (offset == 0 || offset > (size_t)(op - J.out)) { ok = false; break; }
        length = token & 15;
        if (length == 15) {
            uint32_t s;
            do { if (ip >= iend) { ok = false; break; } s = *ip++; length += s; } while (s == 255);
            if (!ok) break;
        }
        length += 4;
        if (length > (size_t)(oend - op) || op + length + 5 > oend) { ok = false; break; }   // last 5 bytes are literals
        __syncwarp();
        if (offset >= 32) {
            for (size_t b = 0; b < length; b += 32) {
                size_t i = b + lane;
                if (i < length) op[i] = op[i - offset];
                __syncwarp();
            }
        } else {
            for (size_t i = lane; i < length; i += 32) op[i] = (op - offset)[i % offset];
            __syncwarp();
        }
        op += length;
    }
done:
    if (!ok && lane == 0) status[J.image] = 0;
}

#include "qoiplane10.cuh"
#include "qoix_sub.cuh"

struct QoiJob { const uint8_t* bytes; uint32_t size; uint8_t* out; uint32_t w, h; int channels; int image; };

__global__ void __launch_bounds__(64)
qoi_kernel(const QoiJob* __restrict__ jobs, int njobs)       // qoi.d:448-550
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= njobs) return;
    const QoiJob J = jobs[j];
    uchar4 index[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) index[i] = make_uchar4(0, 0, 0, 0);
    uchar4 px = make_uchar4(0, 0, 0, 255);
    const uint8_t* b = J.bytes;
    int p = 14, run = 0;
    const int chunks_len = (int)J.size - 8;
    const long long n = (long long)J.w * J.h;
    for (long long i = 0; i < n; ++i) {
        if (run > 0) --run;
        else if (p < chunks_len) {
            const int b1 = b[p++];
            if (b1 == 0xfe) { px.x = b[p++]; px.y = b[p++]; px.z = b[p++]; }
            else if (b1 == 0xff) { px.x = b[p++]; px.y = b[p++]; px.z = b[p++]; px.w = b[p++]; }
            else if ((b1 & 0xc0) == 0x00) px = index[b1];
            else if ((b1 & 0xc0) == 0x40) { px.x += ((b1 >> 4) & 3) - 2; px.y += ((b1 >> 2) & 3) - 2; px.z += (b1 & 3) - 2; }
            else if ((b1 & 0xc0) == 0x80) { const int b2 = b[p++]; const int vg = (b1 & 0x3f) - 32; px.x += vg - 8 + ((b2 >> 4) & 0x0f); px.y += vg; px.z += vg - 8 + (b2 & 0x0f); }
            else run = b1 & 0x3f;
            index[(px.x * 3 + px.y * 5 + px.z * 7 + px.w * 11) & 63] = px;
        }
        if (J.channels == 4) ((uchar4*)J.out)[i] = px;
        else { uint8_t* d = J.out + i * 3; d[0] = px.x; d[1] = px.y; d[2] = px.z; }
    }
}

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
float be32f(const uint8_t* p) { uint32_t v = be32(p); float f; memcpy(&f, &v, 4); return f; }
inline size_t al(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
inline double now_ms() { using namespace std::chrono; return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count(); }

bool valid_load_flags(int f)        // internals/types.d:563-578
{
    if ((f & 0x10000) && (f & 0x80000)) return false;
    if ((f & 0x20000) && (f & 0x40000)) return false;
    if ((f & 0x1000000) && (f & 0x2000000)) return false;
    int n = 0; if (f & 0x100000) ++n; if (f & 0x200000) ++n; if (f & 0x400000) ++n;
    return n <= 1;
}

struct QoixPlan {
    bool ok = false;
    uint32_t w = 0, h = 0; int channels = 0, bitdepth = 0, colorspace = 0, version = 0, compression = 0;
    float par = -1, dpi = -1;
    int type = -1; int codec = -1;       // 0 plane10, 1 qoi10b, 2 plane8, 3 qoi2avg
    uint32_t orig = 0;                    // LZ4: size of the opcode payload
    size_t out_bytes = 0;
};

bool plan_qoix(const uint8_t* d, size_t size, int flags, QoixPlan& P)
{
    if (size < (size_t)QOIX_HEADER_SIZE || size > 0x7fffffffu) return false;
    if (!valid_load_flags(flags)) return false;
    P.version = d[12]; P.channels = d[13]; P.bitdepth = d[14]; P.colorspace = d[15]; P.compression = d[16];
    const bool premul = P.colorspace == 2;
    // identifyTypeFromStream (plugins/qoix.d:476-507)
    if (P.bitdepth == 8) {
        static const int t[5] = {-1, GB200_l8, GB200_la8, GB200_rgb8, GB200_rgba8};
        if (P.channels < 1 || P.channels > 4) return false;
        P.type = t[P.channels]; if (premul && P.channels == 2) P.type = GB200_lap8; if (premul && P.channels == 4) P.type = GB200_rgbap8;
    } else if (P.bitdepth == 10) {
        static const int t[5] = {-1, GB200_l16, GB200_la16, GB200_rgb16, GB200_rgba16};
        if (P.channels < 1 || P.channels > 4) return false;
        P.type = t[P.channels]; if (premul && P.channels == 2) P.type = GB200_lap16; if (premul && P.channels == 4) P.type = GB200_rgbap16;
    } else return false;
    size_t payload = size;
    if (P.compression == 1) {
        if (size < (size_t)QOIX_HEADER_SIZE + 4) return false;
        int orig = (int)be32(d + QOIX_HEADER_SIZE);
        if (orig < 0) return false;
        P.orig = (uint32_t)orig;
        payload = (size_t)QOIX_HEADER_SIZE + P.orig;
    } else if (P.compression != 0) return false;
    if (P.bitdepth == 10) P.codec = ((P.channels == 1 || P.channels == 2) && P.version >= 2) ? 0 : 1;
    else P.codec = (P.channels <= 2) ? 2 : 3;
    // sub-decoder header checks (qoiplane10.d:318-349 and siblings)
    const uint32_t magic = be32(d);
    P.w = be32(d + 4); P.h = be32(d + 8);
    P.par = be32f(d + 17); P.dpi = be32f(d + 21);
    if (P.w == 0 || P.h == 0 || magic != 0x716F6978u || P.h >= 400000000u / P.w) return false;
    if (P.codec == 0) {
        if (payload < (size_t)QOIX_HEADER_SIZE + 5) return false;
        if (P.colorspace > 1 || P.version != 2) return false;      // qoiplane10.d:341-343 (premul streams are rejected)
        P.out_bytes = (size_t)P.w * P.h * P.channels * 2;
    } else if (P.codec == 1) {   // qoi10b.d:510-541
        if (payload < (size_t)QOIX_HEADER_SIZE + 5) return false;
        if (P.colorspace > 2 || P.version > 2) return false;
        P.out_bytes = (size_t)P.w * P.h * P.channels * 2;
    } else if (P.codec == 2) {   // qoiplane.d:379-411
        if (payload < (size_t)QOIX_HEADER_SIZE + 4) return false;
        if (P.colorspace > 1 || P.version > 1) return false;
        P.out_bytes = (size_t)P.w * P.h * P.channels;
    } else {                     // qoi2avg.d:634-665
        if (payload < (size_t)QOIX_HEADER_SIZE + 4) return false;
        if (P.colorspace > 2 || P.version > 1) return false;
        P.out_bytes = (size_t)P.w * P.h * P.channels;
    }
    P.ok = true;
    return true;
}

} // namespace

namespace gb {

gb200_batch* qoix_decode_batch(int n, const uint8_t* const* files, const size_t* lens,
                               const uint8_t* const* files_dev, int flags, cudaStream_t st)
{
    if (!ensure_device()) return nullptr;
    if (n < 0) { set_error("qoix_decode_batch: negative count"); return nullptr; }
    gb200_batch* B = new gb200_batch;
    B->stream = st;
    B->images.resize((size_t)n);
    for (auto& D : B->images) { memset(&D, 0, sizeof(D)); D.ppmX = D.ppmY = D.pixelAspectRatio = -1; }
    double t0 = now_ms();
    std::vector<QoixPlan> P((size_t)n);
    std::vector<int> live;
    size_t out_total = 0, file_total = 0, lz_total = 0;
    std::vector<size_t> out_off((size_t)n, 0), file_off((size_t)n, 0), lz_off((size_t)n, 0);
    for (int i = 0; i < n; ++i) {
        if (!files[i] || !plan_qoix(files[i], lens[i], flags, P[i])) continue;
        live.push_back(i);
        out_off[i] = out_total; out_total += al(P[i].out_bytes);
        file_off[i] = file_total; file_total += al(lens[i] + 16);
        if (P[i].compression == 1) { lz_off[i] = lz_total; lz_total += al((size_t)QOIX_HEADER_SIZE + P[i].orig + 16); }
    }
    B->host_parse_ms = now_ms() - t0;
    const int m = (int)live.size();
    if (!m) return B;
    uint8_t* d_out = (uint8_t*)dev_alloc(out_total);
    if (!d_out) { delete B; return nullptr; }
    B->device_allocs.push_back(d_out);
    DevBuf d_files(files_dev ? 256 : file_total), d_lz(lz_total ? lz_total : 256), d_status(sizeof(int) * (size_t)n);
    if (!d_files.p || !d_lz.p || !d_status.p) { delete B; return nullptr; }
    uint8_t* h_stage = nullptr;
    if (!files_dev) { h_stage = (uint8_t*)pinned_alloc(file_total); if (!h_stage) { delete B; return nullptr; } }
    std::vector<Lz4Job> lz; std::vector<P10Image> pj;
    size_t rec_total = 0, row_total = 0; uint32_t total_chunks = 0;
    std::vector<SubJob> sub[4]; size_t sub_rows_total = 0;
    for (int i : live) {
        const uint8_t* dev = files_dev ? files_dev[i] : d_files.as<uint8_t>() + file_off[i];
        if (!files_dev) memcpy(h_stage + file_off[i], files[i], lens[i]);
        const uint8_t* stream = dev; uint32_t ssize = (uint32_t)lens[i];
        if (P[i].compression == 1) {
            uint8_t* dec = d_lz.as<uint8_t>() + lz_off[i];
            lz.push_back(Lz4Job{dev + QOIX_HEADER_SIZE + 4, (uint32_t)(lens[i] - QOIX_HEADER_SIZE - 4), dec + QOIX_HEADER_SIZE, P[i].orig, i});
            stream = dec; ssize = QOIX_HEADER_SIZE + P[i].orig;
        }
        if (P[i].codec != 0) {
            SubJob S;
            S.stream = stream; S.size = ssize; S.out = d_out + out_off[i]; S.w = P[i].w; S.h = P[i].h;
            S.channels = P[i].channels; S.version = P[i].version; S.image = i;
            S.rows = (uint8_t*)sub_rows_total;                     // offset, rebased below
            sub_rows_total += al((size_t)P[i].w * 2 * (P[i].codec == 1 ? 8 : 4));
            sub[P[i].codec].push_back(S);
            continue;
        }
        P10Image J;
        J.stream = stream; J.size = ssize; J.w = P[i].w; J.h = P[i].h; J.wp = (P[i].w + 7u) & ~7u; J.channels = P[i].channels; J.image = i;
        J.chunk_base = total_chunks; J.nchunks = std::max(1u, (ssize - QOIX_HEADER_SIZE + P10_CHUNK_BYTES - 1) / P10_CHUNK_BYTES);
        J.recs = (uint32_t*)rec_total; J.rowinfo = (uint32_t*)row_total;      // offsets, rebased below
        J.out = d_out + out_off[i];
        total_chunks += J.nchunks; rec_total += al((size_t)J.wp * J.h * 4); row_total += al((size_t)J.h * 4);
        pj.push_back(J);
    }
    DevBuf d_recs(rec_total), d_rows(row_total), d_chunks(sizeof(P10Chunk) * ((size_t)total_chunks + 1)),
           d_entries(sizeof(P10Entry) * ((size_t)total_chunks + 1)), d_dirty(2 * al(total_chunks) + 256), d_misc(256 + 4 * pj.size());
    if (!d_recs.p || !d_rows.p || !d_chunks.p || !d_entries.p || !d_dirty.p || !d_misc.p) { if (h_stage) pinned_free(h_stage); delete B; return nullptr; }
    DevBuf d_subrows(sub_rows_total + 256), d_subjobs(sizeof(SubJob) * (sub[1].size() + sub[2].size() + sub[3].size() + 1));
    if (!d_subrows.p || !d_subjobs.p) { if (h_stage) pinned_free(h_stage); delete B; return nullptr; }
    std::vector<SubJob> suball;
    size_t sub_first[4] = {0, 0, 0, 0};
    for (int c = 1; c < 4; ++c) { sub_first[c] = suball.size(); for (auto& S : sub[c]) { S.rows = d_subrows.as<uint8_t>() + (size_t)S.rows; suball.push_back(S); } }
    for (auto& J : pj) { J.recs = (uint32_t*)(d_recs.as<uint8_t>() + (size_t)J.recs); J.rowinfo = (uint32_t*)(d_rows.as<uint8_t>() + (size_t)J.rowinfo); }
    DevBuf d_lzj(sizeof(Lz4Job) * (lz.size() + 1)), d_pj(sizeof(P10Image) * (pj.size() + 1));
    if (!d_lzj.p || !d_pj.p) { if (h_stage) pinned_free(h_stage); delete B; return nullptr; }
    std::vector<int> ones((size_t)n, 1);
    cudaEvent_t ev[4];
    for (auto& e : ev) cudaEventCreate(&e);
    bool okc = true;
    cudaEventRecord(ev[0], st);
    if (!files_dev) okc &= cuda_ok(cudaMemcpyAsync(d_files.p, h_stage, file_total, cudaMemcpyHostToDevice, st), "files", __FILE__, __LINE__);
    okc &= cuda_ok(cudaMemcpyAsync(d_status.p, ones.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st), "status", __FILE__, __LINE__);
    if (!lz.empty()) okc &= cuda_ok(cudaMemcpyAsync(d_lzj.p, lz.data(), sizeof(Lz4Job) * lz.size(), cudaMemcpyHostToDevice, st), "lz", __FILE__, __LINE__);
    if (!pj.empty()) okc &= cuda_ok(cudaMemcpyAsync(d_pj.p, pj.data(), sizeof(P10Image) * pj.size(), cudaMemcpyHostToDevice, st), "pj", __FILE__, __LINE__);
    if (!suball.empty()) {
        okc &= cuda_ok(cudaMemcpyAsync(d_subjobs.p, suball.data(), sizeof(SubJob) * suball.size(), cudaMemcpyHostToDevice, st), "subjobs", __FILE__, __LINE__);
        okc &= cuda_ok(cudaMemsetAsync(d_subrows.p, 0, sub_rows_total, st), "subrows", __FILE__, __LINE__);
    }
    cudaEventRecord(ev[1], st);
    if (!lz.empty()) { lz4_kernel<<<(unsigned)((lz.size() * 32 + 127) / 128), 128, 0, st>>>(d_lzj.as<Lz4Job>(), (int)lz.size(), d_status.as<int>()); count_launch(); }
    cudaEventRecord(ev[2], st);
    for (int c = 1; c < 4; ++c) {
        if (sub[c].empty()) continue;
        const SubJob* dj = d_subjobs.as<SubJob>() + sub_first[c]; const int nj = (int)sub[c].size();
        if (c == 1) qoi10b_kernel<<<(nj + 31) / 32, 32, 0, st>>>(dj, nj, d_status.as<int>());
        else if (c == 2) qoiplane8_kernel<<<(nj + 31) / 32, 32, 0, st>>>(dj, nj, d_status.as<int>());
        else qoi2avg_kernel<<<(nj + 31) / 32, 32, 0, st>>>(dj, nj, d_status.as<int>());
        count_launch();
    }
    if (!pj.empty()) {
        // QOI-Plane10: chunk-parallel parse (self-synchronising), scan, per-pixel records, wavefront reconstruction
        const P10Image* dI = d_pj.as<P10Image>(); const int ni = (int)pj.size();
        P10Chunk* chunks = d_chunks.as<P10Chunk>(); P10Entry* entries = d_entries.as<P10Entry>();
        uint8_t* dirty[2] = {d_dirty.as<uint8_t>(), d_dirty.as<uint8_t>() + al(total_chunks)};
        uint32_t* changed = d_misc.as<uint32_t>(); uint32_t* ndec = changed + 64;
        const unsigned cg = (total_chunks + 127) / 128;
        p10_sync_kernel<<<cg, 128, 0, st>>>(dI, ni, total_chunks, chunks, dirty[1], dirty[0], 0, changed);
        count_launch();
        for (int pass = 1;; ++pass) {
            uint32_t h_changed = 0;
            okc &= cuda_ok(cudaMemsetAsync(changed, 0, 4, st), "changed", __FILE__, __LINE__);
            p10_sync_kernel<<<cg, 128, 0, st>>>(dI, ni, total_chunks, chunks, dirty[(pass + 1) & 1], dirty[pass & 1], pass, changed);
            count_launch();
            okc &= cuda_ok(cudaMemcpyAsync(&h_changed, changed, 4, cudaMemcpyDeviceToHost, st), "changed back", __FILE__, __LINE__);
            okc &= cuda_ok(cudaStreamSynchronize(st), "sync pass", __FILE__, __LINE__);
            if (!okc || h_changed == 0) break;
        }
        p10_scan_kernel<<<ni, 256, 0, st>>>(dI, chunks, entries, ndec);
        p10_write_kernel<<<cg, 128, 0, st>>>(dI, ni, total_chunks, chunks, entries);
        p10_recon_kernel<1><<<ni, 32, 0, st>>>(dI, ndec, d_status.as<int>());
        p10_recon_kernel<2><<<ni, 32, 0, st>>>(dI, ndec, d_status.as<int>());
        count_launch(4);
    }
    cudaEventRecord(ev[3], st);
    std::vector<int> status((size_t)n);
    okc &= cuda_ok(cudaMemcpyAsync(status.data(), d_status.p, sizeof(int) * n, cudaMemcpyDeviceToHost, st), "status back", __FILE__, __LINE__);
    okc &= cuda_ok(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__);
    okc &= cuda_ok(cudaGetLastError(), "kernels", __FILE__, __LINE__);
    if (h_stage) pinned_free(h_stage);
    if (okc) for (int q = 0; q < 3; ++q) { float ms = 0; cudaEventElapsedTime(&ms, ev[q], ev[q + 1]); B->phase_ms[q] += ms; }
    for (auto& e : ev) cudaEventDestroy(e);
    if (!okc) { delete B; return nullptr; }
    for (int i : live) {
        if (!status[i]) continue;
        gb200_image_desc& D = B->images[i];
        D.status = 1; D.pixels = d_out + out_off[i];
        D.width = (int)P[i].w; D.height = (int)P[i].h; D.channels = P[i].channels; D.file_channels = P[i].channels;
        D.bits = P[i].bitdepth == 10 ? 16 : 8;
        D.pixel_type = P[i].type;
        D.pitch = D.width * D.channels * (D.bits / 8);
        D.pixelAspectRatio = P[i].par; D.ppmY = P[i].dpi;
    }
    B->device_ms = now_ms() - t0 - B->host_parse_ms;
    return B;
}

} // namespace gb

GB_API gb200_batch* gb200_qoix_decode_batch(int n, const uint8_t* const* files, const size_t* lens,
                                            const uint8_t* const* files_dev, int flags, void* stream)
{
    gb::clear_error();
    return gb::qoix_decode_batch(n, files, lens, files_dev, flags, (cudaStream_t)stream);
}

// qoix_lz4_decode (plugins/qoix.d:350): host bytes in, malloc'd host pixels out; *desc filled like the
// sub-decoders do; *decodedType = the stream's own PixelType.
GB_API uint8_t* gb200_qoix_decode(const uint8_t* data, int size, gb200_qoix_desc* desc, int flags, int* decodedType)
{
    gb::clear_error();
    if (!gb::ensure_device()) return nullptr;
    if (size < 0) return nullptr;
    const uint8_t* f[1] = {data}; size_t l[1] = {(size_t)size};
    cudaStream_t st = gb::thread_stream();
    gb200_batch* B = gb::qoix_decode_batch(1, f, l, nullptr, flags, st);
    if (!B) return nullptr;
    const gb200_image_desc& D = B->images[0];
    if (!D.status) { gb::set_error("QOIX decoding failed"); delete B; return nullptr; }
    size_t bytes = (size_t)D.pitch * D.height;
    uint8_t* out = (uint8_t*)malloc(bytes ? bytes : 1);
    if (!out) { delete B; return nullptr; }
    bool ok = gb::cuda_ok(cudaMemcpyAsync(out, D.pixels, bytes, cudaMemcpyDeviceToHost, st), "d2h", __FILE__, __LINE__) &&
              gb::cuda_ok(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__);
    if (desc) {
        desc->width = (uint32_t)D.width; desc->height = (uint32_t)D.height; desc->pitchBytes = D.pitch;
        desc->channels = data[13]; desc->bitdepth = data[14]; desc->colorspace = data[15]; desc->compression = 0;
        desc->pixelAspectRatio = D.pixelAspectRatio; desc->resolutionY = D.ppmY;
    }
    if (decodedType) *decodedType = D.pixel_type;
    delete B;
    if (!ok) { free(out); return nullptr; }
    return out;
}

// qoi_decode (qoi.d:448-550). channels: 0 = as stored, 3 or 4.
GB_API uint8_t* gb200_qoi_decode(const uint8_t* data, int size, gb200_qoi_desc* desc, int channels)
{
    gb::clear_error();
    if (!gb::ensure_device()) return nullptr;
    if ((channels != 0 && channels != 3 && channels != 4) || size < 14 + 8) return nullptr;
    const uint32_t magic = be32(data), w = be32(data + 4), h = be32(data + 8);
    const int fch = data[12], cs = data[13];
    if (desc) { desc->width = w; desc->height = h; desc->channels = (uint8_t)fch; desc->colorspace = (uint8_t)cs; }
    if (w == 0 || h == 0 || fch < 3 || fch > 4 || cs > 1 || magic != 0x716F6966u || h >= 400000000u / w) return nullptr;
    if (channels == 0) channels = fch;
    const size_t bytes = (size_t)w * h * channels;
    cudaStream_t st = gb::thread_stream();
    gb::DevBuf d_in((size_t)size + 16), d_out(bytes), d_job(sizeof(QoiJob));
    if (!d_in.p || !d_out.p || !d_job.p) return nullptr;
    QoiJob J{d_in.as<uint8_t>(), (uint32_t)size, d_out.as<uint8_t>(), w, h, channels, 0};
    uint8_t* out = (uint8_t*)malloc(bytes ? bytes : 1);
    if (!out) return nullptr;
    bool ok = gb::cuda_ok(cudaMemcpyAsync(d_in.p, data, (size_t)size, cudaMemcpyHostToDevice, st), "h2d", __FILE__, __LINE__) &&
              gb::cuda_ok(cudaMemcpyAsync(d_job.p, &J, sizeof(J), cudaMemcpyHostToDevice, st), "job", __FILE__, __LINE__);
    if (ok) { qoi_kernel<<<1, 64, 0, st>>>(d_job.as<QoiJob>(), 1); gb::count_launch(); }
    ok = ok && gb::cuda_ok(cudaGetLastError(), "qoi_kernel", __FILE__, __LINE__) &&
         gb::cuda_ok(cudaMemcpyAsync(out, d_out.p, bytes, cudaMemcpyDeviceToHost, st), "d2h", __FILE__, __LINE__) &&
         gb::cuda_ok(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__);
    if (!ok) { free(out); return nullptr; }
    return out;
}
