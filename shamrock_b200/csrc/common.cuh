// common.cuh — shared device/host helpers for the shamrock_b200 CUDA path (sm_100a).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>

namespace sb {

using u8  = uint8_t;
using u16 = uint16_t;
using u32 = uint32_t;
using u64 = uint64_t;
using i32 = int32_t;
using i64 = int64_t;
using f64 = double;

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define SB_CUDA_CHECK(expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            throw ::sb::CudaError(                                                               \
                std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + __FILE__    \
                + ":" + std::to_string(__LINE__));                                               \
        }                                                                                        \
    } while (0)

#define SB_LAUNCH_CHECK() SB_CUDA_CHECK(cudaGetLastError())

/// number of kernels launched by this library since the last reset (bench.py "gpu_launches")
extern unsigned long long g_launch_count;
#define SB_COUNT_LAUNCH() (++::sb::g_launch_count)

constexpr int kNumSM = 148; // B200

inline unsigned grid_for(u64 n, unsigned block) { return (unsigned) ((n + block - 1) / block); }

/// Caching device allocator (runtime.cu): freed blocks are kept in size buckets and handed out again,
/// so the per-step scratch of the solver never reaches cudaMalloc/cudaFree (which synchronise the
/// device) after the first steps.  Blocks remember the stream they were freed on (runtime.cu) and a
/// different stream waits for that point before it reuses them.  Replaces sham::DeviceBuffer's USM allocations for this path
/// (shambackends/include/shambackends/DeviceBuffer.hpp).
void pool_set_stream(cudaStream_t s); ///< stream the calling thread's allocations / frees are ordered on
void *pool_alloc(size_t bytes);
void pool_free(void *p);
/// Few-word transfers on a compute stream that do not use a copy engine.  A small cudaMemcpyAsync shares the
/// engines with the bulk copies of the host-resident step (Model::evolve_once_host keeps both directions busy for
/// tens of ms) and waits behind them: 8 ms for 3 KB of ghost-zone boxes.  h2d_small passes the bytes as a kernel
/// argument (2 KiB per launch; any host memory, reusable on return); d2h_small stores from a kernel into
/// PAGE-LOCKED host memory (mapped into the device's address space: UVA), complete once the stream is
/// synchronised.  bytes: a multiple of 4, both pointers 4-byte aligned.
void h2d_small(cudaStream_t s, void *d_dst, const void *h_src, size_t bytes);
void d2h_small(cudaStream_t s, void *h_pinned_dst, const void *d_src, size_t bytes);
void pool_release_all();            ///< give every cached block back to the driver
size_t pool_bytes_reserved();

/// grow-only device buffer (no preservation unless asked)
template<class T>
struct DevBuf {
    T *p       = nullptr;
    size_t cap = 0; // elements
    DevBuf()   = default;
    DevBuf(const DevBuf &)            = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), cap(o.cap) {
        o.p   = nullptr;
        o.cap = 0;
    }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) {
            release();
            p     = o.p;
            cap   = o.cap;
            o.p   = nullptr;
            o.cap = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p)
            pool_free(p);
        p   = nullptr;
        cap = 0;
    }
    /// make sure at least n elements fit; contents are lost on growth
    T *ensure(size_t n, double slack = 1.0) {
        if (n > cap) {
            release();
            size_t want = size_t(double(n) * slack) + 16;
            p           = static_cast<T *>(pool_alloc(want * sizeof(T)));
            cap         = want;
        }
        return p;
    }
    /// grow keeping the first `keep` elements
    T *ensure_keep(size_t n, size_t keep, cudaStream_t s, double slack = 1.0) {
        if (n > cap) {
            size_t want = size_t(double(n) * slack) + 16;
            T *np       = static_cast<T *>(pool_alloc(want * sizeof(T)));
            if (p && keep)
                SB_CUDA_CHECK(cudaMemcpyAsync(np, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, s));
            if (p)
                pool_free(p); // stream-ordered reuse: later users are enqueued after the copy
            p   = np;
            cap = want;
        }
        return p;
    }
};

/// pinned host scratch
template<class T>
struct PinnedBuf {
    T *p       = nullptr;
    size_t cap = 0;
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf &)            = delete;
    PinnedBuf &operator=(const PinnedBuf &) = delete;
    PinnedBuf(PinnedBuf &&o) noexcept : p(o.p), cap(o.cap) {
        o.p   = nullptr;
        o.cap = 0;
    }
    PinnedBuf &operator=(PinnedBuf &&o) noexcept {
        if (this != &o) {
            if (p)
                cudaFreeHost(p);
            p     = o.p;
            cap   = o.cap;
            o.p   = nullptr;
            o.cap = 0;
        }
        return *this;
    }
    ~PinnedBuf() {
        if (p)
            cudaFreeHost(p);
    }
    T *ensure(size_t n) {
        if (n > cap) {
            if (p)
                cudaFreeHost(p);
            SB_CUDA_CHECK(cudaMallocHost(&p, n * sizeof(T)));
            cap = n;
        }
        return p;
    }
};

#ifdef __CUDACC__
__device__ __forceinline__ f64 warp_max(f64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ f64 warp_min(f64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ f64 warp_sum(f64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
/// order-preserving map double -> u64 so that atomicMax/Min on u64 orders like the doubles
__device__ __forceinline__ u64 f64_to_ordered(f64 d) {
    u64 b = (u64) __double_as_longlong(d);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ inline f64 ordered_to_f64(u64 k) {
    u64 b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    f64 d;
#ifdef __CUDA_ARCH__
    d = __longlong_as_double((long long) b);
#else
    memcpy(&d, &b, 8);
#endif
    return d;
}
#endif

} // namespace sb
