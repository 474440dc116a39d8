// runtime.cu — process-wide runtime pieces of libshamb200: the caching device allocator and the
// kernel-launch counter.
#include "common.cuh"
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace sb {

unsigned long long g_launch_count = 0;

namespace {
struct Pool {
    std::mutex mu;
    std::multimap<size_t, void *> free_blocks;          // size -> block (per device key folded below)
    std::unordered_map<void *, std::pair<size_t, int>> live; // ptr -> (size, device)
    size_t reserved = 0;
};
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

void *pool_alloc(size_t bytes) {
    Pool &P   = pool();
    size_t sz = round_size(bytes);
    int dev   = 0;
    SB_CUDA_CHECK(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> g(P.mu);
        auto range = P.free_blocks.equal_range(sz);
        for (auto it = range.first; it != range.second; ++it) {
            void *p = it->second;
            if (P.live[p].second == dev) {
                P.free_blocks.erase(it);
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
    P.live[p] = {sz, dev};
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
    P.free_blocks.emplace(it->second.first, p);
}

void pool_release_all() {
    Pool &P = pool();
    std::lock_guard<std::mutex> g(P.mu);
    cudaDeviceSynchronize();
    for (auto &kv : P.free_blocks) {
        auto it = P.live.find(kv.second);
        if (it != P.live.end()) {
            P.reserved -= it->second.first;
            P.live.erase(it);
        }
        cudaFree(kv.second);
    }
    P.free_blocks.clear();
}

size_t pool_bytes_reserved() { return pool().reserved; }

} // namespace sb
