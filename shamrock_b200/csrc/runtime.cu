// runtime.cu — process-wide runtime pieces of libshamb200: the caching device allocator and the
// kernel-launch counter.  The allocator is stream aware: every C-ABI entry point binds its context's
// stream to the calling thread (pool_set_stream); a block freed on one stream and handed to another one
// carries an event the new stream waits for, so two contexts / models on one device never share a block
// that queued work of the other still uses.
#include "common.cuh"
#include <algorithm>
#include <cstring>
#include <map>
#include <stdexcept>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace sb {

unsigned long long g_launch_count = 0;

namespace {
struct Block {
    size_t size = 0;
    int device  = 0;
    cudaStream_t freed_on = nullptr; ///< stream whose queued work may still use the block
    cudaEvent_t freed_ev  = nullptr; ///< recorded on freed_on by pool_free (created on first use)
    bool has_ev           = false;
};
struct Pool {
    std::mutex mu;
    std::multimap<size_t, void *> free_blocks; // size -> block (per device key folded below)
    std::unordered_map<void *, Block> live;    // ptr -> block
    size_t reserved = 0;
};
thread_local cudaStream_t t_stream = nullptr;
thread_local bool t_stream_known   = false;
Pool &pool() {
    static Pool p;
    return p;
}
/// bucket rounding: 512 B granularity below 1 MiB, 1/8-octave steps above (<= 12.5 % slack)
size_t round_size(size_t b) {
    if (b < 512)
        return 512;
    if (b <= (size_t(1) << 20))
        return (b + 511) & ~size_t(511);
    int hi      = 63 - __builtin_clzll(b);
    size_t step = size_t(1) << (hi - 3);
    return (b + step - 1) & ~(step - 1);
}
} // namespace

void pool_set_stream(cudaStream_t s) {
    t_stream       = s;
    t_stream_known = true;
}

void *pool_alloc(size_t bytes) {
    Pool &P   = pool();
    size_t sz = round_size(bytes);
    int dev   = 0;
    SB_CUDA_CHECK(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> g(P.mu);
        // best fit: the smallest cached block that holds the request and is at most twice as large (64 KiB
        // for small requests) — the sizes of per-step scratch (objects that change patch, ghost counts) differ
        // from step to step, and an exact-size cache would go back to cudaMalloc (a device synchronisation) for
        // nearly every one of them
        const size_t limit = std::max<size_t>(2 * sz, size_t(64) << 10);
        for (auto it = P.free_blocks.lower_bound(sz); it != P.free_blocks.end() && it->first <= limit; ++it) {
            void *p  = it->second;
            Block &b = P.live[p];
            if (b.device == dev) {
                P.free_blocks.erase(it);
                // stream-ordered reuse: work queued on the freeing stream may still touch the block; a
                // different stream waits for the event recorded at the free
                if (b.has_ev && !(t_stream_known && b.freed_on == t_stream)) {
                    if (t_stream_known)
                        cudaStreamWaitEvent(t_stream, b.freed_ev, 0);
                    else
                        cudaEventSynchronize(b.freed_ev);
                }
                b.has_ev = false;
                return p;
            }
        }
    }
    void *p       = nullptr;
    cudaError_t e = cudaMalloc(&p, sz);
    if (e != cudaSuccess) { // out of memory: drop the cache and retry once
        cudaGetLastError();
        pool_release_all();
        e = cudaMalloc(&p, sz);
    }
    if (e != cudaSuccess)
        throw CudaError(std::string("cudaMalloc of ") + std::to_string(sz) + " bytes failed: " + cudaGetErrorString(e));
    std::lock_guard<std::mutex> g(P.mu);
    Block b;
    b.size   = sz;
    b.device = dev;
    P.live[p] = b;
    P.reserved += sz;
    return p;
}

void pool_free(void *p) {
    if (!p)
        return;
    Pool &P = pool();
    std::lock_guard<std::mutex> g(P.mu);
    auto it = P.live.find(p);
    if (it == P.live.end()) {
        cudaFree(p);
        return;
    }
    Block &b = it->second;
    if (t_stream_known) {
        if (!b.freed_ev && cudaEventCreateWithFlags(&b.freed_ev, cudaEventDisableTiming) != cudaSuccess)
            b.freed_ev = nullptr;
        if (b.freed_ev && cudaEventRecord(b.freed_ev, t_stream) == cudaSuccess) {
            b.freed_on = t_stream;
            b.has_ev   = true;
        } else {
            cudaGetLastError();
            cudaStreamSynchronize(t_stream);
            b.has_ev = false;
        }
    } else {
        b.has_ev = false; // no stream bound to this thread: the caller synchronised (legacy single-stream use)
    }
    P.free_blocks.emplace(b.size, p);
}

void pool_release_all() {
    Pool &P = pool();
    std::lock_guard<std::mutex> g(P.mu);
    cudaDeviceSynchronize();
    for (auto &kv : P.free_blocks) {
        auto it = P.live.find(kv.second);
        if (it != P.live.end()) {
            P.reserved -= it->second.size;
            if (it->second.freed_ev)
                cudaEventDestroy(it->second.freed_ev);
            P.live.erase(it);
        }
        cudaFree(kv.second);
    }
    P.free_blocks.clear();
}

size_t pool_bytes_reserved() { return pool().reserved; }

// ---- small transfers without a copy engine (common.cuh) ---------------------------------------------------
namespace {
struct SmallBlob {
    u32 w[512];
};
__global__ void h2d_small_kernel(u32 *__restrict__ dst, const SmallBlob b, u32 nw) {
    for (u32 i = threadIdx.x; i < nw; i += blockDim.x)
        dst[i] = b.w[i];
}
__global__ void d2h_small_kernel(u32 *__restrict__ dst, const u32 *__restrict__ src, u32 nw) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += gridDim.x * blockDim.x)
        dst[i] = src[i];
}
} // namespace
void h2d_small(cudaStream_t s, void *d_dst, const void *h_src, size_t bytes) {
    if (bytes % 4 || (reinterpret_cast<uintptr_t>(d_dst) % 4) || (reinterpret_cast<uintptr_t>(h_src) % 4))
        throw std::invalid_argument("h2d_small: 4-byte words only");
    if (bytes > 64 * sizeof(SmallBlob)) { // not small: the copy engine after all
        SB_CUDA_CHECK(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, s));
        return;
    }
    const u32 *src = static_cast<const u32 *>(h_src);
    u32 *dst       = static_cast<u32 *>(d_dst);
    for (size_t w0 = 0, nw = bytes / 4; w0 < nw; w0 += 512) {
        SmallBlob b;
        const u32 n = u32(std::min<size_t>(512, nw - w0));
        std::memcpy(b.w, src + w0, size_t(n) * 4);
        h2d_small_kernel<<<1, 128, 0, s>>>(dst + w0, b, n);
        SB_COUNT_LAUNCH();
    }
    SB_LAUNCH_CHECK();
}
void d2h_small(cudaStream_t s, void *h_pinned_dst, const void *d_src, size_t bytes) {
    if (bytes % 4 || (reinterpret_cast<uintptr_t>(h_pinned_dst) % 4) || (reinterpret_cast<uintptr_t>(d_src) % 4))
        throw std::invalid_argument("d2h_small: 4-byte words only");
    if (!bytes)
        return;
    if (bytes > (1u << 20)) {
        SB_CUDA_CHECK(cudaMemcpyAsync(h_pinned_dst, d_src, bytes, cudaMemcpyDeviceToHost, s));
        return;
    }
    const u32 nw = u32(bytes / 4);
    d2h_small_kernel<<<std::min<u32>(64u, (nw + 255) / 256), 256, 0, s>>>(
        static_cast<u32 *>(h_pinned_dst), static_cast<const u32 *>(d_src), nw);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

} // namespace sb
