// runtime.cu — process-wide runtime pieces of libshamb200: the caching device allocator and the
// kernel-launch counter.  The allocator is stream aware: every C-ABI entry point binds its context's
// stream to the calling thread (pool_set_stream); a block freed on one stream and handed to another one
// carries an event the new stream waits for, so two contexts / models on one device never share a block
// that queued work of the other still uses.
#include "common.cuh"
#include <map>
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
        auto range = P.free_blocks.equal_range(sz);
        for (auto it = range.first; it != range.second; ++it) {
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

} // namespace sb
