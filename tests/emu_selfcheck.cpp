// emu_selfcheck.cpp -- kernels that exercise the primitives of tests/cuda_emu.h one by one; the expected values are the
// documented semantics of the CUDA intrinsics (computed in tests/test_emulator_selfcheck.py). Test infrastructure only.
#include "cuda_emu.h"
#include <string.h>

namespace {

// out[t] = { shfl_up(t, 1), shfl_up(t, 5), ballot(t % 3 == 0), inclusive warp prefix sum of t }
__global__ void k_warp(uint32_t* out)
{
    const uint32_t t = threadIdx.x, lane = t & 31;
    const uint32_t a = __shfl_up_sync(0xffffffffu, t, 1), b = __shfl_up_sync(0xffffffffu, t, 5);
    const uint32_t m = __ballot_sync(0xffffffffu, t % 3 == 0);
    uint32_t inc = t;
    for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (uint32_t)d) inc += n; }
    out[4 * t] = a; out[4 * t + 1] = b; out[4 * t + 2] = m; out[4 * t + 3] = inc;
}

// a CTA-wide reduction through shared memory and __syncthreads; shared atomics; one value per CTA
__global__ void k_cta(uint32_t* out)
{
    __shared__ uint32_t s_sum[256];
    __shared__ uint32_t s_or;
    __shared__ int s_max;
    const uint32_t t = threadIdx.x;
    if (t == 0) { s_or = 0; s_max = -1; }
    s_sum[t] = t + blockIdx.x;
    __syncthreads();
    for (uint32_t step = 128; step > 0; step >>= 1) { if (t < step) s_sum[t] += s_sum[t + step]; __syncthreads(); }
    atomicOr(&s_or, 1u << (t & 31));
    atomicMax(&s_max, (int)(t * 7 % 251));
    __syncthreads();
    if (t == 0) { out[3 * blockIdx.x] = s_sum[0]; out[3 * blockIdx.x + 1] = s_or; out[3 * blockIdx.x + 2] = (uint32_t)s_max; }
}

__global__ void k_bits(const uint32_t* in, uint32_t* out, int n)
{
    const int i = (int)(blockIdx.x * 64 + threadIdx.x);
    if (i >= n) return;
    out[4 * i] = (uint32_t)__clz(in[i]); out[4 * i + 1] = (uint32_t)__ffs((int)in[i]);
    out[4 * i + 2] = __byte_perm(in[i], 0, 0x0123); out[4 * i + 3] = __byte_perm(in[i], ~in[i], 0x7531);
}

}  // namespace

extern "C" void emu_selfcheck(uint32_t* warp_out, uint32_t* cta_out, const uint32_t* bits_in, uint32_t* bits_out, int n)
{
    emu::launch(dim3(1), 96, [&] { k_warp(warp_out); });
    emu::launch(dim3(3), 256, [&] { k_cta(cta_out); });
    emu::launch(dim3((unsigned)((n + 63) / 64)), 64, [&] { k_bits(bits_in, bits_out, n); });
}
