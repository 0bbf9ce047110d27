// common.h -- shared host-side plumbing for the gamut_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#define GB_API extern "C" __attribute__((visibility("default")))

namespace gb {

// Sticky per-thread error string, the C-ABI analogue of Image.error(kStr...) (image.d:1563).
void set_error(const char* fmt, ...);
void clear_error();

// Returns false (and records the error) when a CUDA call failed.
bool cuda_ok(cudaError_t e, const char* what, const char* file, int line);
#define GB_CUDA(x) do { if (!gb::cuda_ok((x), #x, __FILE__, __LINE__)) return 0; } while (0)
#define GB_CUDA_NULL(x) do { if (!gb::cuda_ok((x), #x, __FILE__, __LINE__)) return nullptr; } while (0)
#define GB_CUDA_B(x) do { if (!gb::cuda_ok((x), #x, __FILE__, __LINE__)) return false; } while (0)

// Count of kernels this library launched (bench.py's "gpu_launches").
void count_launch(int n = 1);

// Lazy one-time context init on the current device; fails loudly when there is no GPU.
bool ensure_device();
int  sm_count();
int  device_index();     // current CUDA device (-1 if none); cached state is keyed by it

// Size-bucketed caching device allocator (cudaMalloc is ~100 us; the host-pointer entry points
// are called once per image). Thread-safe.
void* dev_alloc(size_t bytes);
void  dev_free(void* p);
void  dev_trim();

// Pinned staging buffers for the host-pointer entry points.
void* pinned_alloc(size_t bytes);
void  pinned_free(void* p);

// Host-side staging copies (pageable caller memory -> pinned staging) of a batch, spread over a few threads: one
// memcpy thread moves ~10 GB/s, which would otherwise dominate the end-to-end time of a batch decode.
struct HostCopy { void* dst; const void* src; size_t n; };
void host_copy_parallel(const HostCopy* copies, size_t count);

// The library's own non-blocking streams for host-pointer entry points (four per thread; index 0 is
// the default, the others are used to overlap H2D / kernel / D2H of consecutive bands).
cudaStream_t thread_stream(int idx = 0);

struct DevBuf {
    void* p = nullptr;
    DevBuf() {}
    explicit DevBuf(size_t n) { p = dev_alloc(n); }
    ~DevBuf() { if (p) dev_free(p); }
    DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
    bool alloc(size_t n) { if (p) dev_free(p); p = dev_alloc(n); return p != nullptr; }
    template <class T> T* as() const { return (T*)p; }
};

// Grow-only device scratch (used where a launch outlives the call that built its job table).
struct Scratch {
    void* p = nullptr; size_t cap = 0;
    void* get(size_t n) { if (n > cap) { if (p) dev_free(p); p = dev_alloc(n); cap = p ? n : 0; } return p; }
};

} // namespace gb
