// png.cu -- PNG decode orchestration: host chunk walk + device inflate / unfilter / finish.
//
// Drop-in for the codec seam stbi_load_from_callbacks / stbi_load_16_from_callbacks
// (source/gamut/codecs/stbdec.d:713-735) as used by loadPNG (source/gamut/plugins/png.d:44-163).
// The chunk walk (stbi__parse_png_file, stbdec.d:1777-2023) is header logic and stays on the host;
// every per-byte / per-pixel step runs on the GPU:
//   IDAT gather -> inflate (inflate.cuh) -> row unfilter wavefront -> finish (png_kernels.cu).
#include "common.h"
#include "batch.h"
#include "png_kernels.cuh"
#include <vector>
#include <chrono>

namespace gb {
void launch_gather(const void* d_segs, int nsegs, cudaStream_t st);
void launch_unfilter(const UnfilterJob* d_jobs, int njobs, int* d_status, const InflateJob* d_inf, cudaStream_t st, int rowpar_threads);
void launch_finish(const FinishJob* d_jobs, int njobs, uint64_t max_pixels, cudaStream_t st);
struct Segment { const uint8_t* src; uint8_t* dst; uint32_t len; };
}

namespace {

struct Reader {   // memory stream with stb semantics: reads past the end yield 0 (stbdec.d:794-803)
    const uint8_t* p; size_t len, pos;
    uint8_t get8() { return pos < len ? p[pos++] : 0; }
    uint32_t get16() { uint32_t z = get8(); return (z << 8) + get8(); }
    uint32_t get32() { uint32_t z = get16(); return (z << 16) + get16(); }
    void skip(int n) { if (n == 0) return; if (n < 0) { pos = len; return; } if (len - pos < (size_t)n) pos = len; else pos += (size_t)n; }
    bool eof() const { return pos >= len; }
};

struct PngHeader {
    bool ok = false;
    uint32_t w = 0, h = 0;
    int depth = 0, color = 0, interlace = 0;
    int img_n = 0;          // channels stored per pixel in the file rows (1 for paletted)
    int pal_img_n = 0;      // 0, 3 or 4
    bool has_trans = false, is_iphone = false;
    uint8_t tc[3] = {0, 0, 0}; uint16_t tc16[3] = {0, 0, 0};
    uint8_t palette[1024];
    uint32_t pal_len = 0;
    std::vector<std::pair<size_t, uint32_t>> idat;   // (offset in file, length)
    uint32_t ioff = 0;
    float ppmX = -1, ppmY = -1, ratio = -1;
    bool header_only_ok = false;   // result of a SCAN_header walk
    int header_img_n = 0;
};

constexpr uint32_t T(char a, char b, char c, char d) { return ((uint32_t)(uint8_t)a << 24) | ((uint32_t)(uint8_t)b << 16) | ((uint32_t)(uint8_t)c << 8) | (uint32_t)(uint8_t)d; }
const uint8_t kDepthScale[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};

// Chunk walk with the accept/reject behaviour of stbi__parse_png_file (stbdec.d:1777-2023).
// scan_header = true reproduces STBI__SCAN_header (used by stbi__png_is16).
bool parse_png(const uint8_t* data, size_t len, bool scan_header, PngHeader& H)
{
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    Reader s{data, len, 0};
    memset(H.palette, 0, sizeof(H.palette));
    for (int i = 0; i < 8; ++i) if (s.get8() != sig[i]) return false;
    bool first = true;
    for (;;) {
        uint32_t clen = s.get32();
        uint32_t ctype = s.get32();
        switch (ctype) {
        case T('C','g','B','I'): H.is_iphone = true; s.skip((int)clen); break;
        case T('p','H','Y','s'): {
            H.ppmX = (float)s.get32(); H.ppmY = (float)s.get32();
            H.ratio = H.ppmX / H.ppmY;
            if (s.get8() != 1) { H.ppmX = -1; H.ppmY = -1; }
            break; }
        case T('I','H','D','R'): {
            if (!first) return false;
            first = false;
            if (clen != 13) return false;
            H.w = s.get32(); H.h = s.get32();
            if (H.h > (1u << 24) || H.w > (1u << 24)) return false;
            H.depth = s.get8();
            if (H.depth != 1 && H.depth != 2 && H.depth != 4 && H.depth != 8 && H.depth != 16) return false;
            H.color = s.get8(); if (H.color > 6) return false;
            if (H.color == 3 && H.depth == 16) return false;
            if (H.color == 3) H.pal_img_n = 3; else if (H.color & 1) return false;
            if (s.get8()) return false;      // compression method
            if (s.get8()) return false;      // filter method
            H.interlace = s.get8(); if (H.interlace > 1) return false;
            if (!H.w || !H.h) return false;
            if (!H.pal_img_n) {
                H.img_n = (H.color & 2 ? 3 : 1) + (H.color & 4 ? 1 : 0);
                if ((1u << 30) / H.w / H.img_n < H.h) return false;
                if (scan_header) { H.header_img_n = H.img_n; return true; }
            } else {
                H.img_n = 1;
                if ((1u << 30) / H.w / 4 < H.h) return false;
            }
            break; }
        case T('P','L','T','E'): {
            if (first) return false;
            if (clen > 256 * 3) return false;
            H.pal_len = clen / 3;
            if (H.pal_len * 3 != clen) return false;
            for (uint32_t i = 0; i < H.pal_len; ++i) {
                H.palette[i*4+0] = s.get8(); H.palette[i*4+1] = s.get8(); H.palette[i*4+2] = s.get8(); H.palette[i*4+3] = 255;
            }
            break; }
        case T('t','R','N','S'): {
            if (first) return false;
            if (H.ioff) return false;                              // z.idata != null <=> some IDAT bytes seen
            if (H.pal_img_n) {
                if (scan_header) { H.header_img_n = 4; return true; }
                if (H.pal_len == 0) return false;
                if (clen > H.pal_len) return false;
                H.pal_img_n = 4;
                for (uint32_t i = 0; i < clen; ++i) H.palette[i*4+3] = s.get8();
            } else {
                if (!(H.img_n & 1)) return false;
                if (clen != (uint32_t)H.img_n * 2) return false;
                H.has_trans = true;
                if (H.depth == 16) { for (int k = 0; k < H.img_n; ++k) H.tc16[k] = (uint16_t)s.get16(); }
                else { for (int k = 0; k < H.img_n; ++k) H.tc[k] = (uint8_t)((uint8_t)(s.get16() & 255) * kDepthScale[H.depth]); }
            }
            break; }
        case T('I','D','A','T'): {
            if (first) return false;
            if (H.pal_img_n && !H.pal_len) return false;
            if (scan_header) { H.header_img_n = H.pal_img_n; return true; }
            if ((int)(H.ioff + clen) < (int)H.ioff) return false;
            if (s.len - s.pos < (size_t)clen) return false;       // stbi__getn fails: "outofdata"
            // (a zero-length IDAT leaves z.idata null in the reference: nothing is recorded)
            if (clen) H.idat.push_back({s.pos, clen});
            s.pos += clen;
            H.ioff += clen;
            break; }
        case T('I','E','N','D'):
            if (first) return false;
            if (scan_header) return true;
            H.ok = !H.idat.empty();
            return H.ok;
        default:
            if (first) return false;
            if (ctype == 0 && s.eof()) {                          // gamut issue #92: no IEND
                if (scan_header) return true;
                H.ok = !H.idat.empty();
                return H.ok;
            }
            if ((ctype & (1u << 29)) == 0) return false;
            s.skip((int)clen);
            break;
        }
        s.get32();   // CRC (not checked)
    }
}

struct Plan {    // everything derived from the header for one image
    PngHeader H;
    bool ok = false;
    int out_n = 0;          // img_out_n at unfilter time
    int cur_n = 0;          // img_out_n after palette / tRNS
    int req_n = 0;          // final channels
    int file_n = 0;         // reported `comp`
    int out16 = 0;
    bool direct = false;    // unfilter writes the final buffer
    int npass = 0;
    uint32_t pass_w[7], pass_h[7], pass_rb[7], pass_raw_off[7], pass_packed_off[7];
    uint32_t need_len = 0;  // inflated bytes needed (sum of (rb+1)*h over passes)
    uint32_t packed_len = 0;
    uint32_t guess = 0, cap = 0;
    size_t out_bytes = 0;
};

bool make_plan(const uint8_t* data, size_t len, int req_comp, int want16, Plan& P)
{
    PngHeader& H = P.H;
    if (req_comp < 0 || req_comp > 4) return false;
    if (!parse_png(data, len, false, H)) return false;
    // finalize_decode (stbdec.d:1800-1858)
    uint32_t bpl = (H.w * H.depth + 7) / 8;
    P.guess = bpl * H.h * H.img_n + H.h;
    if ((req_comp == H.img_n + 1 && req_comp != 3 && !H.pal_img_n) || H.has_trans) P.out_n = H.img_n + 1;
    else P.out_n = H.img_n;
    P.file_n = H.img_n;
    P.cur_n = P.out_n;
    if (H.pal_img_n) {
        P.file_n = H.pal_img_n;
        P.cur_n = H.pal_img_n;
        if (req_comp >= 3) P.cur_n = req_comp;
    } else if (H.has_trans) {
        P.file_n = H.img_n + 1;
    }
    P.req_n = req_comp ? req_comp : P.cur_n;
    int file16 = H.depth == 16;
    P.out16 = want16 < 0 ? file16 : want16;
    // passes
    static const int xorig[7] = {0,4,0,2,0,1,0}, yorig[7] = {0,0,4,0,2,0,1}, xspc[7] = {8,8,4,4,2,2,1}, yspc[7] = {8,8,8,4,4,2,2};
    uint32_t raw_off = 0, packed_off = 0;
    P.npass = H.interlace ? 7 : 1;
    for (int p = 0; p < P.npass; ++p) {
        uint32_t x = H.w, y = H.h;
        if (H.interlace) { x = (H.w - xorig[p] + xspc[p] - 1) / xspc[p]; y = (H.h - yorig[p] + yspc[p] - 1) / yspc[p]; }
        P.pass_w[p] = x; P.pass_h[p] = y;
        if (!x || !y) { P.pass_w[p] = P.pass_h[p] = 0; P.pass_rb[p] = 0; P.pass_raw_off[p] = raw_off; P.pass_packed_off[p] = packed_off; continue; }
        uint32_t rb = ((H.img_n * x * H.depth) + 7) >> 3;
        if (H.depth < 8 && rb > x) return false;          // "invalid width" (stbdec.d:1442)
        P.pass_rb[p] = rb;
        P.pass_raw_off[p] = raw_off; P.pass_packed_off[p] = packed_off;
        raw_off += (rb + 1) * y;
        packed_off += rb * y;
    }
    P.need_len = raw_off;
    P.packed_len = packed_off;
    P.out_bytes = (size_t)H.w * H.h * P.req_n * (P.out16 ? 2 : 1);
    P.direct = (H.depth == 8 && !H.interlace && !H.pal_img_n && !H.has_trans && P.out_n == H.img_n &&
                P.req_n == P.cur_n && !P.out16);
    // first buffer size of the grow sequence that can hold the stream (stbdec.d:1296-1309)
    P.cap = P.guess;
    P.ok = true;
    return true;
}

inline size_t al(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
inline double now_ms() { using namespace std::chrono; return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count(); }

} // namespace

namespace gb {

// Decodes a batch. files_dev may be null (then the IDAT payloads are staged through pinned memory).
gb200_batch* png_decode_batch(int n, const uint8_t* const* files, const size_t* lens,
                              const uint8_t* const* files_dev, int req_comp, int want16, cudaStream_t st)
{
    if (!ensure_device()) return nullptr;
    if (n < 0) { set_error("png_decode_batch: negative count"); return nullptr; }
    gb200_batch* B = new gb200_batch;
    B->stream = st;
    B->images.resize((size_t)n);
    double t0 = now_ms();
    std::vector<Plan> plans((size_t)n);
    size_t out_total = 0;
    std::vector<size_t> out_off((size_t)n, 0);
    for (int i = 0; i < n; ++i) {
        gb200_image_desc& D = B->images[i];
        memset(&D, 0, sizeof(D));
        D.ppmX = D.ppmY = D.pixelAspectRatio = -1;
        Plan& P = plans[i];
        bool ok = files[i] && make_plan(files[i], lens[i], req_comp, want16, P);
        D.ppmX = P.H.ppmX; D.ppmY = P.H.ppmY; D.pixelAspectRatio = P.H.ratio;
        if (!ok) { P.ok = false; continue; }
        out_off[i] = out_total;
        out_total += al(P.out_bytes);
    }
    B->host_parse_ms = now_ms() - t0;

    uint8_t* d_out = nullptr;
    if (out_total) {
        d_out = (uint8_t*)dev_alloc(out_total);
        if (!d_out) { delete B; return nullptr; }
        B->device_allocs.push_back(d_out);
    }
    std::vector<int> pending;
    for (int i = 0; i < n; ++i) if (plans[i].ok) pending.push_back(i);
    std::vector<int> final_ok((size_t)n, 0);

    while (!pending.empty()) {
        const int m = (int)pending.size();
        // ---- layout of this round's scratch
        size_t idat_total = 0, raw_total = 0, packed_total = 0;
        std::vector<size_t> idat_off(m), raw_off(m), packed_off(m);
        int nseg = 0, nunf = 0, nfin = 0;
        for (int k = 0; k < m; ++k) {
            Plan& P = plans[pending[k]];
            idat_off[k] = idat_total; idat_total += al(P.H.ioff + 32);
            raw_off[k] = raw_total;   raw_total += al((size_t)P.cap + 32);
            if (!P.direct) { packed_off[k] = packed_total; packed_total += al((size_t)P.packed_len + 32); nfin++; }
            nseg += (int)P.H.idat.size();
            for (int p = 0; p < P.npass; ++p) if (P.pass_w[p]) nunf++;
        }
        DevBuf d_idat(idat_total), d_raw(raw_total), d_packed(packed_total ? packed_total : 256);
        DevBuf d_status(sizeof(int) * (size_t)m);
        if (!d_idat.p || !d_raw.p || !d_packed.p || !d_status.p) { delete B; return nullptr; }
        std::vector<InflateJob> ijobs(m);
        std::vector<UnfilterJob> ujobs; ujobs.reserve(nunf);
        std::vector<FinishJob> fjobs; fjobs.reserve(nfin);
        std::vector<Segment> segs; segs.reserve(nseg);
        uint8_t* h_stage = nullptr;
        if (!files_dev) { h_stage = (uint8_t*)pinned_alloc(idat_total); if (!h_stage) { delete B; return nullptr; } }
        uint64_t max_pixels = 1;
        int rowpar_threads = 32;
        std::vector<HostCopy> hcopies;
        for (int k = 0; k < m; ++k) {
            const int i = pending[k];
            Plan& P = plans[i];
            uint8_t* idat = d_idat.as<uint8_t>() + idat_off[k];
            size_t o = 0;
            for (auto& sg : P.H.idat) {
                if (files_dev) segs.push_back(Segment{files_dev[i] + sg.first, idat + o, sg.second});
                else hcopies.push_back(HostCopy{h_stage + idat_off[k] + o, files[i] + sg.first, sg.second});
                o += sg.second;
            }
            if (!files_dev) memset(h_stage + idat_off[k] + o, 0, al(P.H.ioff + 32) - o);
            InflateJob& ij = ijobs[k];
            ij.in = idat; ij.in_len = P.H.ioff;
            ij.out = d_raw.as<uint8_t>() + raw_off[k]; ij.out_cap = P.cap;
            ij.parse_header = P.H.is_iphone ? 0 : 1;
            ij.out_len = 0; ij.status = 0;
            uint8_t* outp = d_out + out_off[i];
            uint8_t* packed = P.direct ? outp : d_packed.as<uint8_t>() + packed_off[k];
            int bpp = P.H.depth < 8 ? 1 : P.H.img_n * (P.H.depth == 16 ? 2 : 1);
            for (int p = 0; p < P.npass; ++p) {
                if (!P.pass_w[p]) continue;
                UnfilterJob u;
                u.raw = ij.out + P.pass_raw_off[p];
                u.out = packed + P.pass_packed_off[p];
                u.row_bytes = P.pass_rb[p]; u.height = P.pass_h[p]; u.bpp = (uint32_t)bpp; u.out_pitch = P.pass_rb[p];
                u.image = k; u.inflate_idx = k; u.need_len = P.need_len;
                ujobs.push_back(u);
                if (bpp == 4) { int t = (int)((P.pass_rb[p] / 4 + 127) / 128) * 32; if (t <= 1024 && t > rowpar_threads) rowpar_threads = t; }
            }
            if (!P.direct) {
                FinishJob f; memset(&f, 0, sizeof(f));
                f.packed = packed;
                for (int p = 0; p < 7; ++p) {
                    f.pass_off[p] = p < P.npass ? P.pass_packed_off[p] : 0;
                    f.pass_w[p] = p < P.npass ? P.pass_w[p] : 0; f.pass_h[p] = p < P.npass ? P.pass_h[p] : 0;
                    f.pass_rb[p] = p < P.npass ? P.pass_rb[p] : 0;
                }
                f.out = outp; f.w = P.H.w; f.h = P.H.h;
                f.depth = (uint8_t)P.H.depth; f.color = (uint8_t)P.H.color; f.interlace = (uint8_t)P.H.interlace; f.img_n = (uint8_t)P.H.img_n;
                f.add_alpha = P.out_n != P.H.img_n; f.has_trans = P.H.has_trans; f.pal_n = P.H.pal_img_n ? (uint8_t)P.cur_n : 0;
                f.cur_n = (uint8_t)P.cur_n; f.req_n = (uint8_t)P.req_n; f.out16 = (uint8_t)P.out16;
                memcpy(f.tc, P.H.tc, 3); memcpy(f.tc16, P.H.tc16, 6); memcpy(f.palette, P.H.palette, 1024);
                fjobs.push_back(f);
                uint64_t px = (uint64_t)P.H.w * P.H.h; if (px > max_pixels) max_pixels = px;
            }
        }
        host_copy_parallel(hcopies.data(), hcopies.size());
        DevBuf d_ij(sizeof(InflateJob) * (size_t)m), d_uj(sizeof(UnfilterJob) * (ujobs.size() + 1)),
               d_fj(sizeof(FinishJob) * (fjobs.size() + 1)), d_sg(sizeof(Segment) * (segs.size() + 1));
        if (!d_ij.p || !d_uj.p || !d_fj.p || !d_sg.p) { if (h_stage) pinned_free(h_stage); delete B; return nullptr; }
        std::vector<int> ones((size_t)m, 1);
        bool okc = true;
        okc &= cuda_ok(cudaMemcpyAsync(d_status.p, ones.data(), sizeof(int) * m, cudaMemcpyHostToDevice, st), "status", __FILE__, __LINE__);
        okc &= cuda_ok(cudaMemcpyAsync(d_ij.p, ijobs.data(), sizeof(InflateJob) * m, cudaMemcpyHostToDevice, st), "ijobs", __FILE__, __LINE__);
        if (!ujobs.empty()) okc &= cuda_ok(cudaMemcpyAsync(d_uj.p, ujobs.data(), sizeof(UnfilterJob) * ujobs.size(), cudaMemcpyHostToDevice, st), "ujobs", __FILE__, __LINE__);
        if (!fjobs.empty()) okc &= cuda_ok(cudaMemcpyAsync(d_fj.p, fjobs.data(), sizeof(FinishJob) * fjobs.size(), cudaMemcpyHostToDevice, st), "fjobs", __FILE__, __LINE__);
        cudaEvent_t ev[5];
        for (auto& e : ev) cudaEventCreate(&e);
        cudaEventRecord(ev[0], st);
        if (files_dev) {
            okc &= dev_fill_async(d_idat.p, 0, idat_total, st);
            if (!segs.empty()) okc &= cuda_ok(cudaMemcpyAsync(d_sg.p, segs.data(), sizeof(Segment) * segs.size(), cudaMemcpyHostToDevice, st), "segs", __FILE__, __LINE__);
            launch_gather(d_sg.p, (int)segs.size(), st);
        } else {
            okc &= cuda_ok(cudaMemcpyAsync(d_idat.p, h_stage, idat_total, cudaMemcpyHostToDevice, st), "idat", __FILE__, __LINE__);
        }
        cudaEventRecord(ev[1], st);
        InflateWork iwork;
        okc &= launch_inflate(d_ij.as<InflateJob>(), ijobs.data(), m, st, iwork);
        cudaEventRecord(ev[2], st);
        launch_unfilter(d_uj.as<UnfilterJob>(), (int)ujobs.size(), d_status.as<int>(), d_ij.as<InflateJob>(), st, rowpar_threads);
        cudaEventRecord(ev[3], st);
        launch_finish(d_fj.as<FinishJob>(), (int)fjobs.size(), max_pixels, st);
        cudaEventRecord(ev[4], st);
        std::vector<int> status((size_t)m);
        {
            // results come back through pinned memory written by a kernel, not through the copy engine (common.h)
            const size_t ij_bytes = (sizeof(InflateJob) * (size_t)m + 15) & ~(size_t)15;
            PinnedBuf h_back(ij_bytes + sizeof(int) * (size_t)m);
            if (!h_back.p) okc = false;
            okc = okc && dev_read_back_async(h_back.p, d_ij.p, sizeof(InflateJob) * m, st);
            okc = okc && dev_read_back_async(h_back.as<uint8_t>() + ij_bytes, d_status.p, sizeof(int) * m, st);
            okc &= cuda_ok(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__);
            if (okc) { memcpy(ijobs.data(), h_back.p, sizeof(InflateJob) * m); memcpy(status.data(), h_back.as<uint8_t>() + ij_bytes, sizeof(int) * m); }
        }
        okc &= cuda_ok(cudaGetLastError(), "kernels", __FILE__, __LINE__);
        if (h_stage) pinned_free(h_stage);
        if (okc) for (int q = 0; q < 4; ++q) { float ms = 0; cudaEventElapsedTime(&ms, ev[q], ev[q + 1]); B->phase_ms[q] += ms; }
        for (auto& e : ev) cudaEventDestroy(e);
        if (!okc) { cudaStreamSynchronize(st); delete B; return nullptr; }     // nothing in flight may outlive the scratch it uses

        std::vector<int> next;
        for (int k = 0; k < m; ++k) {
            const int i = pending[k];
            Plan& P = plans[i];
            if (ijobs[k].status == INF_OUTPUT_FULL) {
                // MZ_BUF_ERROR: grow like stbdec.d:1296-1309 and decode again from scratch
                if (P.cap > 536870912u) continue;              // fail
                uint64_t c = (uint64_t)P.cap * 2;
                if (c < 32 * 1024) c = 32 * 1024;
                if (c > 0xffffffffull - 64) continue;
                P.cap = (uint32_t)c;
                next.push_back(i);
                continue;
            }
            if (ijobs[k].status != INF_OK) continue;
            if (ijobs[k].out_len < P.need_len) continue;         // "not enough pixels" (stbdec.d:1430)
            if (!status[k]) continue;                            // invalid filter byte
            final_ok[i] = 1;
        }
        pending.swap(next);
    }

    for (int i = 0; i < n; ++i) {
        gb200_image_desc& D = B->images[i];
        Plan& P = plans[i];
        if (!final_ok[i]) { D.status = 0; D.pixels = nullptr; continue; }
        D.status = 1;
        D.pixels = d_out + out_off[i];
        D.width = (int)P.H.w; D.height = (int)P.H.h;
        D.channels = P.req_n; D.file_channels = P.file_n;
        D.bits = P.out16 ? 16 : 8;
        static const int t8[5] = {-1, GB200_l8, GB200_la8, GB200_rgb8, GB200_rgba8};
        static const int t16[5] = {-1, GB200_l16, GB200_la16, GB200_rgb16, GB200_rgba16};
        D.pixel_type = P.out16 ? t16[P.req_n] : t8[P.req_n];     // plugins/png.d:120-157
        D.pitch = D.width * D.channels * (P.out16 ? 2 : 1);
    }
    B->device_ms = now_ms() - t0 - B->host_parse_ms;
    return B;
}

} // namespace gb

// ---------------------------------------------------------------------------------------------
// C ABI
GB_API int gb200_batch_count(const gb200_batch* b) { return b ? (int)b->images.size() : 0; }
GB_API const gb200_image_desc* gb200_batch_images(const gb200_batch* b) { return b ? b->images.data() : nullptr; }
GB_API void gb200_batch_free(gb200_batch* b) { delete b; }
GB_API void gb200_batch_timing(const gb200_batch* b, float* phase_ms8, double* host_parse_ms)
{
    if (!b) return;
    if (phase_ms8) for (int i = 0; i < 8; ++i) phase_ms8[i] = b->phase_ms[i];
    if (host_parse_ms) *host_parse_ms = b->host_parse_ms;
}

GB_API int gb200_batch_download(const gb200_batch* b, uint8_t* dst_host, size_t stride)
{
    gb::clear_error();
    if (!b) return 0;
    if (!gb::ensure_device()) return 0;
    // one asynchronous copy per decoded image on the batch's stream, one synchronisation for all of them
    // (dst_host should be pinned -- gb200_host_alloc -- for the copies to run at PCIe speed)
    for (size_t i = 0; i < b->images.size(); ++i) {
        const gb200_image_desc& D = b->images[i];
        if (!D.status || !D.pixels) continue;
        GB_CUDA(cudaMemcpyAsync(dst_host + i * stride, D.pixels, (size_t)D.pitch * D.height, cudaMemcpyDeviceToHost, b->stream));
    }
    GB_CUDA(cudaStreamSynchronize(b->stream));
    return 1;
}

GB_API int gb200_png_is16(const uint8_t* data, size_t len)
{
    PngHeader H;
    if (!data || !parse_png(data, len, true, H)) return 0;
    return H.depth == 16;
}

GB_API gb200_batch* gb200_png_decode_batch(int n, const uint8_t* const* files, const size_t* lens,
                                           const uint8_t* const* files_dev, int req_comp, int want16, void* stream)
{
    gb::clear_error();
    return gb::png_decode_batch(n, files, lens, files_dev, req_comp, want16, (cudaStream_t)stream);
}

GB_API uint8_t* gb200_png_load(const uint8_t* data, size_t len, int req_comp, int want16,
                               int* width, int* height, int* comp, float* ppmX, float* ppmY, float* pixelRatio)
{
    gb::clear_error();
    if (ppmX) *ppmX = -1; if (ppmY) *ppmY = -1; if (pixelRatio) *pixelRatio = -1;
    if (!gb::ensure_device()) return nullptr;
    const uint8_t* f[1] = {data}; size_t l[1] = {len};
    cudaStream_t st = gb::thread_stream();
    gb200_batch* B = gb::png_decode_batch(1, f, l, nullptr, req_comp, want16 ? 1 : 0, st);
    if (!B) return nullptr;
    const gb200_image_desc& D = B->images[0];
    if (ppmX) *ppmX = D.ppmX; if (ppmY) *ppmY = D.ppmY; if (pixelRatio) *pixelRatio = D.pixelAspectRatio;
    if (!D.status) { gb::set_error("PNG decoding failed"); delete B; return nullptr; }
    size_t bytes = (size_t)D.pitch * D.height;
    uint8_t* out = (uint8_t*)malloc(bytes ? bytes : 1);
    if (!out) { delete B; return nullptr; }
    bool ok = gb::cuda_ok(cudaMemcpyAsync(out, D.pixels, bytes, cudaMemcpyDeviceToHost, st), "d2h", __FILE__, __LINE__) &&
              gb::cuda_ok(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__);
    if (width) *width = D.width; if (height) *height = D.height; if (comp) *comp = D.file_channels;
    delete B;
    if (!ok) { free(out); return nullptr; }
    return out;
}

GB_API int gb200_png_unfilter_device(const uint8_t* raw, size_t raw_stride, uint8_t* out, size_t out_stride,
                                     int n_images, int row_bytes, int height, int bpp, int* status_dev, void* stream)
{
    gb::clear_error();
    if (!gb::ensure_device()) return 0;
    if (n_images <= 0 || row_bytes <= 0 || height <= 0) return 1;
    if (!(bpp == 1 || bpp == 2 || bpp == 3 || bpp == 4 || bpp == 6 || bpp == 8) || row_bytes % bpp) { gb::set_error("unfilter: bad bpp"); return 0; }
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<gb::UnfilterJob> jobs((size_t)n_images);
    for (int i = 0; i < n_images; ++i) {
        gb::UnfilterJob& u = jobs[i];
        u.raw = raw + (size_t)i * raw_stride; u.out = out + (size_t)i * out_stride;
        u.row_bytes = (uint32_t)row_bytes; u.height = (uint32_t)height; u.bpp = (uint32_t)bpp; u.out_pitch = (uint32_t)row_bytes;
        u.image = i; u.inflate_idx = -1; u.need_len = 0;
    }
    // persistent per-thread scratch: the launch is asynchronous, so the job table must outlive the call
    //  the previous launch that read the table may still be running, possibly on another stream: wait for it
    // before the table is overwritten or regrown
    static thread_local gb::Scratch s_jobs, s_dummy;
    static thread_local cudaEvent_t s_done = nullptr;
    if (s_done) GB_CUDA(cudaEventSynchronize(s_done));
    else GB_CUDA(cudaEventCreateWithFlags(&s_done, cudaEventDisableTiming));
    gb::UnfilterJob* d_jobs = (gb::UnfilterJob*)s_jobs.get(sizeof(gb::UnfilterJob) * (size_t)n_images);
    int* d_dummy = (int*)s_dummy.get(sizeof(int) * (size_t)n_images);
    if (!d_jobs || !d_dummy) return 0;
    GB_CUDA(cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(gb::UnfilterJob) * (size_t)n_images, cudaMemcpyHostToDevice, st));
    int rp = bpp == 4 ? (int)(((row_bytes / 4 + 127) / 128) * 32) : 32;
    gb::launch_unfilter(d_jobs, n_images, status_dev ? status_dev : d_dummy, nullptr, st, rp <= 1024 ? rp : 32);
    GB_CUDA(cudaGetLastError());
    GB_CUDA(cudaEventRecord(s_done, st));
    // jobs were copied from pageable memory: the copy has completed on return
    return 1;
}

GB_API void gb200_inflate_set_mode(int parallel) { gb::set_inflate_mode(parallel ? 1 : 0); }

GB_API int gb200_inflate_device(int n, const uint8_t* const* in_dev, const uint32_t* in_lens,
                                uint8_t* const* out_dev, const uint32_t* out_caps, int parse_header,
                                uint32_t* out_lens_dev, int* statuses_dev, void* stream)
{
    gb::clear_error();
    if (!gb::ensure_device()) return 0;
    if (n <= 0) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<gb::InflateJob> jobs((size_t)n);
    for (int i = 0; i < n; ++i) {
        jobs[i].in = in_dev[i]; jobs[i].in_len = in_lens[i]; jobs[i].out = out_dev[i]; jobs[i].out_cap = out_caps[i];
        jobs[i].parse_header = parse_header; jobs[i].out_len = 0; jobs[i].status = 0;
        if ((uintptr_t)in_dev[i] & 3) { gb::set_error("inflate: input streams must be 4-byte aligned"); return 0; }
    }
    gb::DevBuf d_jobs(sizeof(gb::InflateJob) * (size_t)n);
    if (!d_jobs.p) return 0;
    GB_CUDA(cudaMemcpyAsync(d_jobs.p, jobs.data(), sizeof(gb::InflateJob) * (size_t)n, cudaMemcpyHostToDevice, st));
    gb::InflateWork iwork;
    if (!gb::launch_inflate(d_jobs.as<gb::InflateJob>(), jobs.data(), n, st, iwork)) return 0;
    GB_CUDA(cudaGetLastError());
    if (out_lens_dev) GB_CUDA(cudaMemcpy2DAsync(out_lens_dev, 4, (const char*)d_jobs.p + offsetof(gb::InflateJob, out_len), sizeof(gb::InflateJob), 4, n, cudaMemcpyDeviceToDevice, st));
    if (statuses_dev) GB_CUDA(cudaMemcpy2DAsync(statuses_dev, 4, (const char*)d_jobs.p + offsetof(gb::InflateJob, status), sizeof(gb::InflateJob), 4, n, cudaMemcpyDeviceToDevice, st));
    GB_CUDA(cudaStreamSynchronize(st));
    return 1;
}
