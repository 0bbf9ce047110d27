// bmp_encode.cuh -- the header and the kernel of the BMP writer (host code in bmp_encode.cu). Compiled for the GPU by
// bmp_encode.cu and for the host, under the thread-per-CUDA-thread emulation, by tests/emu_bmp_encode.cpp.
//
// Reference: saveBMP (plugins/bmp.d:166-194) -> write_bmp (codecs/bmpenc.d:25-113): a 122-byte header (BITMAPFILEHEADER
// + a 108-byte V4 DIB header; BI_BITFIELDS with B / G / R / A masks for 32-bit files), then the scanlines bottom-up,
// RGB -> BGR / RGBA -> BGRA (scanline.d:812-834), each padded to a multiple of 4 bytes. The reference writes the padding
// from an uninitialised malloc() buffer (:40-43, :103): those bytes are not defined by it; this writer stores zeros.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <math.h>

namespace {

constexpr int BE_HEADER = 14 + 108;

struct BeImage {
    const uint8_t* pixels; int pitch;         // first scanline, signed pitch; rgb8 / rgba8
    int w, h, channels, pad;
    uint8_t* out;                             // header + h rows of (w * channels + pad) bytes
    uint8_t header[BE_HEADER];
};

// one thread per pixel of the file; the thread of a row's last pixel also writes the row's padding
__global__ void __launch_bounds__(256)
be_rows_kernel(const BeImage* __restrict__ imgs)
{
    const BeImage& im = imgs[blockIdx.z];
    const int x = (int)(blockIdx.x * 256u + threadIdx.x), y = (int)blockIdx.y;       // y = row of the file, bottom row first
    if (y >= im.h) return;
    if (y == 0 && blockIdx.x == 0 && threadIdx.x < BE_HEADER) im.out[threadIdx.x] = im.header[threadIdx.x];
    if (x >= im.w) return;
    const uint8_t* src = im.pixels + (ptrdiff_t)im.pitch * (ptrdiff_t)(im.h - 1 - y) + (size_t)x * im.channels;
    uint8_t* dst = im.out + BE_HEADER + (size_t)y * ((size_t)im.w * im.channels + im.pad) + (size_t)x * im.channels;
    dst[0] = src[2]; dst[1] = src[1]; dst[2] = src[0];
    if (im.channels == 4) dst[3] = src[3];
    if (x == im.w - 1) for (int k = 0; k < im.pad; ++k) dst[im.channels + k] = 0;
}

// ---- host side (shared with the emulation harness) -----------------------------------------------------------------
// type = PixelType value: rgb8 = 9, rgba8 = 12 (saveBMP, plugins/bmp.d:174-183); sides 1..32767 (:188-189)
inline int be_channels(int type) { return type == 9 ? 3 : type == 12 ? 4 : 0; }
inline size_t be_size(int type, int width, int height)
{
    const int ch = be_channels(type);
    if (!ch || width < 1 || height < 1 || width > 32767 || height > 32767) return 0;
    const int linesize = width * ch, pad = 3 - ((linesize - 1) & 3);
    return (size_t)BE_HEADER + (size_t)height * (size_t)(linesize + pad);
}
// ppmX / ppmY: Image.pixelsPerMeterX / Y (image.d:344-361), -1 (GAMUT_UNKNOWN_RESOLUTION) when unknown
inline bool be_setup(BeImage& B, const uint8_t* pixels, int type, int width, int height, int pitch, float ppmX, float ppmY, uint8_t* out)
{
    const size_t filesize = be_size(type, width, height);
    if (!filesize || !pixels || !out || filesize > 0xffffffffull) return false;
    const int ch = be_channels(type);
    const long long ap = pitch < 0 ? -(long long)pitch : pitch;
    if (ap < (long long)width * ch && height > 1) return false;
    B = BeImage();
    B.pixels = pixels; B.pitch = pitch; B.w = width; B.h = height; B.channels = ch; B.pad = 3 - ((width * ch - 1) & 3);
    B.out = out;
    uint8_t* h = B.header;                                         // bmpenc.d:45-90; everything not set below is 0
    auto le32 = [&](int at, uint32_t v) { h[at] = (uint8_t)v; h[at + 1] = (uint8_t)(v >> 8); h[at + 2] = (uint8_t)(v >> 16); h[at + 3] = (uint8_t)(v >> 24); };
    h[0] = 0x42; h[1] = 0x4d;
    le32(2, (uint32_t)filesize);
    le32(10, BE_HEADER);
    le32(14, 108);
    le32(18, (uint32_t)width); le32(22, (uint32_t)height);         // positive height: bottom-up
    h[26] = 1; h[28] = (uint8_t)(ch * 8);
    le32(30, ch == 3 ? 0u : 3u);                                   // CMP_RGB / CMP_BITS
    int ippmX = 0, ippmY = 0;
    if (ppmX != -1.0f) ippmX = (int)round((double)ppmX);
    if (ppmY != -1.0f) ippmY = (int)round((double)ppmY);
    le32(38, (uint32_t)ippmX); le32(42, (uint32_t)ippmY);
    if (ch == 4) { h[56] = 0xff; h[59] = 0xff; h[62] = 0xff; h[69] = 0xff; }      // R, G, B, A masks (:73-79)
    h[70] = 'B'; h[71] = 'G'; h[72] = 'R'; h[73] = 's';
    return true;
}

}  // namespace
