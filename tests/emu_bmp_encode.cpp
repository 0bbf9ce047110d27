// emu_bmp_encode.cpp -- gamut_b200/csrc/bmp_encode.cuh compiled for the host under tests/cuda_emu.h; the launch is the one
// of gb200_bmp_encode (bmp_encode.cu). Test infrastructure only.
#include "cuda_emu.h"
#include "../gamut_b200/csrc/bmp_encode.cuh"

extern "C" long emu_bmp_encode(const uint8_t* pixels, int type, int width, int height, int pitch, float ppmX, float ppmY, uint8_t* out, size_t out_cap)
{
    const size_t filesize = be_size(type, width, height);
    if (!filesize) return 0;
    if (filesize > out_cap) return -1;
    BeImage B;
    if (!be_setup(B, pixels, type, width, height, pitch, ppmX, ppmY, out)) return 0;
    const BeImage* dI = &B;
    emu::launch(dim3((unsigned)((width + 255) / 256), (unsigned)height, 1), 256, [&] { be_rows_kernel(dI); });
    return (long)filesize;
}
