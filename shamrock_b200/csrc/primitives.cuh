// primitives.cuh — CUB-free device-wide primitives: exclusive scan, min/max reductions, fills.
// replaces (for this path) shamalgs::numeric::scan_exclusive (decoupled look-back,
// shamalgs/src/details/numeric/numeric.cpp:48-59), stream_compact
// (details/numeric/streamCompactExclScan.cpp:96) and the group reductions
// (details/reduction/groupReduction_usm.cpp).
#pragma once
#include "common.cuh"

namespace sb {

// ---------------------------------------------------------------------------------------------
// exclusive scan: reduce-then-scan, 3 launches; 2 reads + 1 write of the data.
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS   = 8;
constexpr int SCAN_TILE    = SCAN_THREADS * SCAN_ITEMS; // 2048

template<class Tin>
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(
    const Tin *__restrict__ in, u64 n, u32 *__restrict__ block_sums) {
    __shared__ u32 warp_s[SCAN_THREADS / 32];
    u64 base = u64(blockIdx.x) * SCAN_TILE;
    u32 s    = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        u64 i = base + u64(k) * SCAN_THREADS + threadIdx.x;
        if (i < n)
            s += u32(in[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0)
        warp_s[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 t = 0;
#pragma unroll
        for (int w = 0; w < SCAN_THREADS / 32; w++)
            t += warp_s[w];
        block_sums[blockIdx.x] = t;
    }
}

/// single block: exclusive scan of block_sums in place; total -> *total (u64 to detect overflow)
static __global__ void __launch_bounds__(1024) scan_sums_kernel(u32 *block_sums, u32 nblocks, u64 *total) {
    __shared__ u64 warp_s[32];
    __shared__ u64 carry_s;
    if (threadIdx.x == 0)
        carry_s = 0;
    __syncthreads();
    for (u32 base = 0; base < nblocks; base += 1024) {
        u32 i   = base + threadIdx.x;
        u64 v   = (i < nblocks) ? u64(block_sums[i]) : 0;
        u64 inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u64 t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((threadIdx.x & 31) >= o)
                inc += t;
        }
        if ((threadIdx.x & 31) == 31)
            warp_s[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            u64 w  = warp_s[threadIdx.x];
            u64 wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                u64 t = __shfl_up_sync(0xffffffffu, wi, o);
                if (threadIdx.x >= o)
                    wi += t;
            }
            warp_s[threadIdx.x] = wi - w; // exclusive warp offsets
        }
        __syncthreads();
        u64 excl = carry_s + warp_s[threadIdx.x >> 5] + (inc - v);
        if (i < nblocks)
            block_sums[i] = u32(excl);
        __syncthreads();
        if (threadIdx.x == 1023)
            carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        *total = carry_s;
}

template<class Tin>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(
    const Tin *__restrict__ in, u64 n, const u32 *__restrict__ block_sums, u32 *__restrict__ out) {
    __shared__ u32 warp_s[SCAN_THREADS / 32];
    // blocked arrangement: thread t owns items [t*ITEMS, t*ITEMS+ITEMS) of the tile
    u64 base = u64(blockIdx.x) * SCAN_TILE + u64(threadIdx.x) * SCAN_ITEMS;
    u32 v[SCAN_ITEMS];
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        u64 i = base + k;
        v[k]  = (i < n) ? u32(in[i]) : 0u;
        s += v[k];
    }
    u32 inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((threadIdx.x & 31) >= o)
            inc += t;
    }
    if ((threadIdx.x & 31) == 31)
        warp_s[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
        u32 w  = (threadIdx.x < SCAN_THREADS / 32) ? warp_s[threadIdx.x] : 0u;
        u32 wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, wi, o);
            if (threadIdx.x >= o)
                wi += t;
        }
        if (threadIdx.x < SCAN_THREADS / 32)
            warp_s[threadIdx.x] = wi - w;
    }
    __syncthreads();
    u32 run = block_sums[blockIdx.x] + warp_s[threadIdx.x >> 5] + (inc - s);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        u64 i = base + k;
        if (i < n)
            out[i] = run;
        run += v[k];
    }
}

/// out[i] = sum_{j<i} in[j]; *d_total = sum (u64).  tmp must hold ceil(n/2048) u32.
template<class Tin>
inline void exclusive_scan(
    cudaStream_t s, const Tin *in, u32 *out, u64 n, DevBuf<u32> &tmp, u64 *d_total) {
    if (n == 0) {
        SB_CUDA_CHECK(cudaMemsetAsync(d_total, 0, sizeof(u64), s));
        return;
    }
    u32 nb = u32((n + SCAN_TILE - 1) / SCAN_TILE);
    tmp.ensure(nb);
    scan_block_sums_kernel<Tin><<<nb, SCAN_THREADS, 0, s>>>(in, n, tmp.p);
    SB_COUNT_LAUNCH();
    scan_sums_kernel<<<1, 1024, 0, s>>>(tmp.p, nb, d_total);
    SB_COUNT_LAUNCH();
    scan_apply_kernel<Tin><<<nb, SCAN_THREADS, 0, s>>>(in, n, tmp.p, out);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
template<class T>
__global__ void fill_kernel(T *p, u64 n, T v) {
    u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
        p[i] = v;
}
template<class T>
inline void fill(cudaStream_t s, T *p, u64 n, T v) {
    if (!n)
        return;
    fill_kernel<T><<<grid_for(n, 256), 256, 0, s>>>(p, n, v);
    SB_COUNT_LAUNCH();
}

} // namespace sb
