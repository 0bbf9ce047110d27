// cuda_emu.h -- a thread-per-CUDA-thread emulation of the few CUDA constructs the encoder kernels use, so that the
// text of a .cuh can be compiled with g++ and checked against the oracle on a machine without a GPU (test
// infrastructure only; nothing in gamut_b200/ includes it). One CTA runs at a time: 'blockDim' host threads execute
// the kernel body, __syncthreads() is a barrier over them, warp primitives are a barrier over the 32 threads of a
// warp around a slot array, __shared__ is a function-local static (shared by the threads of the CTA that is running).
// Limits: kernels whose threads leave at different barriers, and anything that relies on the hardware's memory model
// beyond barriers and atomics, are not modelled.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <pthread.h>
#include <algorithm>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct emu_uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static thread_local emu_uint3 threadIdx, blockIdx;
static emu_uint3 blockDim, gridDim;

namespace emu {
static pthread_barrier_t cta_barrier;
static pthread_barrier_t warp_barrier[32];
static uint32_t warp_slot[32][32];
}

static inline void __syncthreads() { pthread_barrier_wait(&emu::cta_barrier); }

template <class T> static inline T __shfl_up_sync(unsigned, T v, int d)
{
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t raw; __builtin_memcpy(&raw, &v, 4);
    emu::warp_slot[warp][lane] = raw;
    pthread_barrier_wait(&emu::warp_barrier[warp]);
    uint32_t got = (int)lane >= d ? emu::warp_slot[warp][lane - d] : raw;
    pthread_barrier_wait(&emu::warp_barrier[warp]);
    T r; __builtin_memcpy(&r, &got, 4);
    return r;
}
static inline uint32_t __ballot_sync(unsigned, bool p)
{
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    emu::warp_slot[warp][lane] = p ? 1u : 0u;
    pthread_barrier_wait(&emu::warp_barrier[warp]);
    uint32_t m = 0;
    for (int l = 0; l < 32; ++l) m |= emu::warp_slot[warp][l] << l;
    pthread_barrier_wait(&emu::warp_barrier[warp]);
    return m;
}
static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&emu::warp_barrier[threadIdx.x >> 5]); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int atomicMax(int* a, int v) { int old = __atomic_load_n(a, __ATOMIC_RELAXED); while (old < v && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {} return old; }
static inline uint32_t atomicOr(uint32_t* a, uint32_t v) { return __atomic_fetch_or(a, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T __ldg(const T* p) { return *p; }
struct uint4 { uint32_t x, y, z, w; };
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel)
{
    const uint64_t v = (uint64_t)b << 32 | a;
    uint32_t r = 0;
    for (int k = 0; k < 4; ++k) r |= (uint32_t)((v >> (8 * ((sel >> (4 * k)) & 7))) & 0xff) << (8 * k);
    return r;
}
static inline int __clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
using std::max;
using std::min;

namespace emu {
// kernel<<<grid, threads>>>(args...)  ->  emu::launch(grid, threads, [&] { kernel(args...); })
static inline void launch(dim3 grid, unsigned threads, const std::function<void()>& body)
{
    blockDim = {threads, 1, 1}; gridDim = {grid.x, grid.y, grid.z};
    const unsigned warps = (threads + 31) / 32;
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
            pthread_barrier_init(&cta_barrier, nullptr, threads);
            for (unsigned w = 0; w < warps; ++w) pthread_barrier_init(&warp_barrier[w], nullptr, std::min(32u, threads - w * 32));
            std::vector<std::thread> ts;
            ts.reserve(threads);
            for (unsigned t = 0; t < threads; ++t)
                ts.emplace_back([&, t] { threadIdx = {t, 0, 0}; blockIdx = {bx, by, 0}; body(); });
            for (auto& th : ts) th.join();
            pthread_barrier_destroy(&cta_barrier);
            for (unsigned w = 0; w < warps; ++w) pthread_barrier_destroy(&warp_barrier[w]);
        }
}
}
