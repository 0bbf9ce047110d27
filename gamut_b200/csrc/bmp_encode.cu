// bmp_encode.cu -- BMP writer on the GPU: saveBMP (plugins/bmp.d:166-194) -> write_bmp
// (codecs/bmpenc.d:25-113). Header, kernel and their description: bmp_encode.cuh.
#include "../../include/gamut_b200.h"
#include "common.h"
#include "bmp_encode.cuh"
#include <cstring>

GB_API size_t gb200_bmp_encode_size(const gb200_bmp_desc* desc)
{
    return desc ? be_size(desc->type, desc->width, desc->height) : 0;
}

// `pixels` = the first scanline of an rgb8 / rgba8 image on the host, desc->pitchBytes signed. malloc()'d file out (free
// with gb200_free), *out_len its length; NULL where saveBMP fails (other pixel types, a side outside 1..32767).
GB_API uint8_t* gb200_bmp_encode(const uint8_t* pixels, const gb200_bmp_desc* desc, int* out_len)
{
    gb::clear_error();
    if (!gb::ensure_device()) return nullptr;
    const size_t filesize = desc ? be_size(desc->type, desc->width, desc->height) : 0;
    if (!pixels || !out_len || !filesize || filesize > 0x7fffffffull) {
        gb::set_error("bmp_encode: unsupported image (BMP takes rgb8 / rgba8 with sides 1..32767)");
        return nullptr;
    }
    const int ch = be_channels(desc->type);
    const size_t row = (size_t)desc->width * ch;
    const size_t ap = desc->pitchBytes < 0 ? (size_t)(-(long long)desc->pitchBytes) : (size_t)desc->pitchBytes;
    if (ap < row && desc->height > 1) { gb::set_error("bmp_encode: pitch smaller than a scanline"); return nullptr; }
    cudaStream_t st = gb::thread_stream();
    const size_t span = ap * (size_t)(desc->height - 1) + row;
    const uint8_t* lowest = desc->pitchBytes < 0 ? pixels - ap * (size_t)(desc->height - 1) : pixels;
    gb::DevBuf d_in(span), d_out(filesize), d_img(sizeof(BeImage));
    if (!d_in.p || !d_out.p || !d_img.p) return nullptr;
    BeImage B;
    if (!be_setup(B, d_in.as<uint8_t>() + (pixels - lowest), desc->type, desc->width, desc->height, desc->pitchBytes, desc->ppmX, desc->ppmY,
                  d_out.as<uint8_t>())) return nullptr;
    uint8_t* out = (uint8_t*)malloc(filesize);
    if (!out) return nullptr;
    bool ok = gb::cuda_ok(cudaMemcpyAsync(d_in.p, lowest, span, cudaMemcpyHostToDevice, st), "be h2d", __FILE__, __LINE__) &&
              gb::cuda_ok(cudaMemcpyAsync(d_img.p, &B, sizeof(B), cudaMemcpyHostToDevice, st), "be img", __FILE__, __LINE__);
    if (ok) {
        be_rows_kernel<<<dim3((unsigned)((desc->width + 255) / 256), (unsigned)desc->height, 1), 256, 0, st>>>(d_img.as<BeImage>());
        gb::count_launch();
        ok = gb::cuda_ok(cudaMemcpyAsync(out, d_out.p, filesize, cudaMemcpyDeviceToHost, st), "be d2h", __FILE__, __LINE__);
    }
    ok = gb::cuda_ok(cudaStreamSynchronize(st), "be sync", __FILE__, __LINE__) && ok;       // B and the buffers outlive the work
    ok = ok && gb::cuda_ok(cudaGetLastError(), "be kernel", __FILE__, __LINE__);
    if (!ok) { free(out); return nullptr; }
    *out_len = (int)filesize;
    return out;
}
