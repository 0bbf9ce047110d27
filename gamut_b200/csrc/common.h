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

// Fill and small read-back WITHOUT the copy engines. On this platform a cudaMemsetAsync or a device-to-host
// cudaMemcpyAsync queues behind every download already issued on ANY stream (one FIFO per engine; measured: a 64 KB
// read-back issued after a 1 GiB download starts when that download ends), which serialised the decode of sub-batch
// k+1 behind the download of sub-batch k in gb200_decode_batch_host. These two run as kernels: dev_fill_async sets n
// bytes, dev_read_back_async copies n bytes of device memory into PINNED host memory (mapped under UVA) with a store
// from the SMs; the caller synchronises the stream before reading it.
bool dev_fill_async(void* p, int byte, size_t n, cudaStream_t st);
bool dev_read_back_async(void* pinned_dst, const void* dev_src, size_t n, cudaStream_t st);

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

struct PinnedBuf {
    void* p = nullptr;
    explicit PinnedBuf(size_t n) { p = pinned_alloc(n ? n : 1); }
    ~PinnedBuf() { if (p) pinned_free(p); }
    PinnedBuf(const PinnedBuf&) = delete; PinnedBuf& operator=(const PinnedBuf&) = delete;
    template <class T> T* as() const { return (T*)p; }
};

// Grow-only device scratch (used where a launch outlives the call that built its job table).
struct Scratch {
    void* p = nullptr; size_t cap = 0;
    void* get(size_t n) { if (n > cap) { if (p) dev_free(p); p = dev_alloc(n); cap = p ? n : 0; } return p; }
};

} // namespace gb
