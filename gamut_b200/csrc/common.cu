// common.cu -- error state, device init, caching allocator, launch counter.
#include "common.h"
#include <thread>
#include <vector>
#include <stdarg.h>
#include <mutex>
#include <map>
#include <vector>
#include <atomic>
#include <array>

namespace gb {


static thread_local char t_err[512] = {0};
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...)
{
    va_list ap; va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}
void clear_error() { t_err[0] = 0; }

bool cuda_ok(cudaError_t e, const char* what, const char* file, int line)
{
    if (e == cudaSuccess) return true;
    set_error("CUDA error %s (%s) at %s:%d in %s", cudaGetErrorName(e), cudaGetErrorString(e), file, line, what);
    return false;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Per-device state: a process may drive several devices (cudaSetDevice between calls); everything cached is keyed by
// the current device.
constexpr int MAX_DEV = 64;
static std::mutex g_dev_m;
static int g_dev_state[MAX_DEV];       // 0 unknown, 1 usable, -1 unusable
static int g_sms[MAX_DEV];

static int current_device() { int d = 0; if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); return -1; } return d; }

bool ensure_device()
{
    const int dev = current_device();
    if (dev < 0 || dev >= MAX_DEV) {
        set_error("gamut_b200: no CUDA device available; there is no CPU fallback");
        return false;
    }
    {
        std::lock_guard<std::mutex> g(g_dev_m);
        if (g_dev_state[dev] == 0) {
            cudaDeviceProp p;
            if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { cudaGetLastError(); g_dev_state[dev] = -1; }
            else if (p.major != 10) { set_error("gamut_b200: built for sm_100a only, found sm_%d%d", p.major, p.minor); g_dev_state[dev] = -1; }
            else { g_sms[dev] = p.multiProcessorCount; g_dev_state[dev] = 1; }
        }
        if (g_dev_state[dev] == 1) return true;
    }
    if (t_err[0] == 0) set_error("gamut_b200: no usable sm_100 CUDA device; there is no CPU fallback");
    return false;
}
int sm_count() { const int d = current_device(); return d >= 0 && d < MAX_DEV && g_sms[d] > 0 ? g_sms[d] : 148; }
int device_index() { return current_device(); }

// ---------------------------------------------------------------------------------------------
// Caching allocator: power-of-two-ish buckets (round up to 1/8 octave above 1 MiB, 512 B below).
// Keyed per device.
struct Pool {
    std::mutex m;
    std::multimap<std::pair<int, size_t>, void*> free_;   // (device,size) -> ptr
    std::map<void*, std::pair<int, size_t>> live_;
    size_t cached_bytes = 0;
};
static Pool& dpool() { static Pool* p = new Pool; return *p; }
static Pool& hpool() { static Pool* p = new Pool; return *p; }

static size_t bucket(size_t n)
{
    if (n < 512) return 512;
    size_t p = 512;
    while (p < n) p <<= 1;
    // 8 sub-buckets per octave to bound waste at 12.5 %
    size_t step = p >> 4;
    size_t lo = p >> 1;
    size_t b = lo + ((n - lo + step - 1) / step) * step;
    return b < n ? p : b;
}

static void* pool_alloc(Pool& P, size_t bytes, bool pinned)
{
    if (!ensure_device()) return nullptr;
    int dev = 0; cudaGetDevice(&dev);
    if (pinned) dev = -1;
    size_t b = bucket(bytes ? bytes : 1);
    {
        std::lock_guard<std::mutex> g(P.m);
        auto it = P.free_.find({dev, b});
        if (it != P.free_.end()) {
            void* p = it->second;
            P.free_.erase(it);
            P.cached_bytes -= b;
            P.live_[p] = {dev, b};
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = pinned ? cudaMallocHost(&p, b) : cudaMalloc(&p, b);
    if (e != cudaSuccess) {
        // release the cache and retry once
        cudaGetLastError();
        if (pinned) { /* nothing */ } else dev_trim();
        e = pinned ? cudaMallocHost(&p, b) : cudaMalloc(&p, b);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("gamut_b200: out of %s memory allocating %zu bytes", pinned ? "pinned host" : "device", b);
            return nullptr;
        }
    }
    std::lock_guard<std::mutex> g(P.m);
    P.live_[p] = {dev, b};
    return p;
}
// Cached (free) bytes are capped: beyond the cap a block goes back to the driver instead of the cache. The caps are
// generous for the batch workloads (their scratch is reused call after call) and bound what an idle process holds.
static const size_t DEV_CACHE_CAP = (size_t)64 << 30, PINNED_CACHE_CAP = (size_t)16 << 30;
static void pool_free(Pool& P, void* p, bool pinned)
{
    if (!p) return;
    bool release = false;
    {
        std::lock_guard<std::mutex> g(P.m);
        auto it = P.live_.find(p);
        if (it == P.live_.end()) return;
        if (P.cached_bytes + it->second.second > (pinned ? PINNED_CACHE_CAP : DEV_CACHE_CAP)) release = true;
        else { P.free_.insert({it->second, p}); P.cached_bytes += it->second.second; }
        P.live_.erase(it);
    }
    if (release) { if (pinned) cudaFreeHost(p); else cudaFree(p); }
}

void* dev_alloc(size_t bytes) { return pool_alloc(dpool(), bytes, false); }
void  dev_free(void* p) { pool_free(dpool(), p, false); }
void  dev_trim()
{
    Pool& P = dpool();
    std::vector<void*> v;
    {
        std::lock_guard<std::mutex> g(P.m);
        for (auto& kv : P.free_) v.push_back(kv.second);
        P.free_.clear();
        P.cached_bytes = 0;
    }
    for (void* p : v) cudaFree(p);
}
void* pinned_alloc(size_t bytes) { return pool_alloc(hpool(), bytes, true); }
void  pinned_free(void* p) { pool_free(hpool(), p, true); }

cudaStream_t thread_stream(int idx)
{
    // eight streams per (thread, device): a stream belongs to the device that was current when it was created
    static thread_local std::map<int, std::array<cudaStream_t, 8>> per_dev;
    idx &= 7;
    if (!ensure_device()) return nullptr;
    const int dev = current_device();
    auto it = per_dev.find(dev);
    if (it == per_dev.end()) it = per_dev.emplace(dev, std::array<cudaStream_t, 8>{}).first;
    cudaStream_t& st = it->second[idx];
    if (!st && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); st = nullptr; }
    return st;
}

__global__ void __launch_bounds__(256) fill_kernel(uint8_t* p, size_t n, uint32_t v4)
{
    // head bytes up to the first 16-byte boundary, 16-byte vectors, tail bytes
    const size_t head = min(n, (size_t)((16 - ((uintptr_t)p & 15)) & 15));
    const size_t nvec = (n - head) >> 4, tail0 = head + (nvec << 4);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    if (tid < head) p[tid] = (uint8_t)v4;
    uint4* v = (uint4*)(p + head);
    for (size_t i = tid; i < nvec; i += nth) v[i] = make_uint4(v4, v4, v4, v4);
    if (tid < n - tail0) p[tail0 + tid] = (uint8_t)v4;
}
__global__ void __launch_bounds__(256) read_back_kernel(uint8_t* dst, const uint8_t* src, size_t n)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    if ((((uintptr_t)dst | (uintptr_t)src) & 3) == 0) {
        const size_t nw = n >> 2;
        for (size_t i = tid; i < nw; i += nth) ((uint32_t*)dst)[i] = ((const uint32_t*)src)[i];
        if (tid < (n & 3)) dst[(nw << 2) + tid] = src[(nw << 2) + tid];
    } else for (size_t i = tid; i < n; i += nth) dst[i] = src[i];
}
bool dev_fill_async(void* p, int byte, size_t n, cudaStream_t st)
{
    if (n == 0) return true;
    const uint32_t b = (uint32_t)byte & 255u;
    size_t blocks = (n / 16 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > (size_t)sm_count() * 16) blocks = (size_t)sm_count() * 16;
    fill_kernel<<<(unsigned)blocks, 256, 0, st>>>((uint8_t*)p, n, b * 0x01010101u);
    count_launch();
    return cuda_ok(cudaGetLastError(), "fill_kernel", __FILE__, __LINE__);
}
bool dev_read_back_async(void* pinned_dst, const void* dev_src, size_t n, cudaStream_t st)
{
    if (n == 0) return true;
    size_t blocks = (n / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 64) blocks = 64;
    read_back_kernel<<<(unsigned)blocks, 256, 0, st>>>((uint8_t*)pinned_dst, (const uint8_t*)dev_src, n);
    count_launch();
    return cuda_ok(cudaGetLastError(), "read_back_kernel", __FILE__, __LINE__);
}

void host_copy_parallel(const HostCopy* copies, size_t count)
{
    size_t total = 0;
    for (size_t i = 0; i < count; ++i) total += copies[i].n;
    unsigned hw = std::thread::hardware_concurrency();
    unsigned nt = total < (8u << 20) ? 1u : (hw >= 16 ? 8u : hw >= 4 ? hw / 2 : 1u);
    if (nt <= 1) { for (size_t i = 0; i < count; ++i) memcpy(copies[i].dst, copies[i].src, copies[i].n); return; }
    // thread t takes the byte range [t, t+1) * total / nt of the concatenated copies
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) {
        th.emplace_back([=]() {
            const size_t lo = total * t / nt, hi = total * (t + 1) / nt;
            size_t pos = 0;
            for (size_t i = 0; i < count && pos < hi; ++i) {
                const size_t a = pos, b = pos + copies[i].n;
                pos = b;
                if (b <= lo) continue;
                const size_t s0 = a < lo ? lo - a : 0, s1 = (b > hi ? hi : b) - a;
                if (s1 > s0) memcpy((char*)copies[i].dst + s0, (const char*)copies[i].src + s0, s1 - s0);
            }
        });
    }
    for (auto& x : th) x.join();
}

} // namespace gb

// ---------------------------------------------------------------------------------------------
// C ABI: library-level entry points (declared in include/gamut_b200.h)
GB_API const char* gb200_last_error(void) { return gb::t_err; }
GB_API long long gb200_launch_count(void) { return gb::g_launches.load(); }
GB_API int gb200_init(void) { return gb::ensure_device() ? 1 : 0; }
GB_API int gb200_sm_count(void) { return gb::ensure_device() ? gb::sm_count() : 0; }
GB_API void* gb200_device_alloc(size_t bytes) { return gb::dev_alloc(bytes); }
GB_API void gb200_device_free(void* p) { gb::dev_free(p); }
GB_API void gb200_device_trim(void) { gb::dev_trim(); }
GB_API void* gb200_host_alloc(size_t bytes) { return gb::pinned_alloc(bytes); }
GB_API void gb200_host_free(void* p) { gb::pinned_free(p); }
GB_API const char* gb200_version(void) { return "gamut_b200 0.1 (sm_100a)"; }
GB_API void gb200_free(void* p) { free(p); }
GB_API int gb200_copy_to_host(void* dst_host, const void* src_dev, size_t bytes)
{
    if (!gb::ensure_device()) return 0;
    GB_CUDA(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
    return 1;
}
/* Download by the SMs instead of a copy engine: a grid-stride 16-byte copy whose stores land in PINNED host memory
 * (mapped under UVA). Asynchronous on `stream`. dst_pinned, src_dev and bytes must be multiples of 16. */
namespace gb {
__global__ void __launch_bounds__(256) sm_download_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = __ldcs(src + i);
}
}
GB_API int gb200_download_by_kernel(void* dst_pinned, const void* src_dev, size_t bytes, void* stream)
{
    if (!gb::ensure_device()) return 0;
    if ((((uintptr_t)dst_pinned | (uintptr_t)src_dev | bytes) & 15) != 0) { gb::set_error("gb200_download_by_kernel: 16-byte alignment"); return 0; }
    if (!bytes) return 1;
    gb::sm_download_kernel<<<gb::sm_count() * 8, 256, 0, (cudaStream_t)stream>>>((uint4*)dst_pinned, (const uint4*)src_dev, bytes / 16);
    gb::count_launch();
    GB_CUDA(cudaGetLastError());
    return 1;
}
GB_API int gb200_copy_to_device(void* dst_dev, const void* src_host, size_t bytes)
{
    if (!gb::ensure_device()) return 0;
    GB_CUDA(cudaMemcpy(dst_dev, src_host, bytes, cudaMemcpyHostToDevice));
    return 1;
}
