// microbench.cu — roofline denominators measured on the device the context lives on: FP64 FMA-chain
// throughput and a streaming copy.  Pattern of the reference's own device micro-benchmarks
// (shamsys/src/MicroBenchmark.cpp:51-77, shambackends/include/shambackends/benchmarks/fma_chains.hpp,
// saxpy): the FP64 roof of the SPH loops is not part of MEASURED_PEAKS.json, so bench.py measures it.
#include "solver.cuh"

namespace sb {

template<int CHAINS>
__global__ void __launch_bounds__(256) fma_chain_kernel(f64 *out, int iters, f64 a, f64 b) {
    f64 x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++)
        x[c] = f64(threadIdx.x + c) * 1e-3;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++)
            x[c] = fma(x[c], a, b);
    }
    f64 s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++)
        s += x[c];
    if (s == 12345.678) // never true: keeps the chains alive
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) stream_copy_kernel(const double2 *__restrict__ in, double2 *__restrict__ out, u64 n) {
    for (u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += u64(gridDim.x) * blockDim.x)
        out[i] = in[i];
}

// ---- 32-byte record gathers (the access pattern of the SPH neighbour loops) ----------------------------
// One 256-bit load per lane and trip from a table of 2^k records; the record index is a function of a
// per-warp counter and the lane, so no index traffic competes with the gather.  PATTERN:
//   0  32 consecutive records per warp (coalesced: the floor, 8 data wavefronts per instruction)
//   1  an independent random record per lane (what a neighbour list looks like)
//   2  a random 128-byte line per lane, bank group (address bits 5-6) = lane & 3
//   3  a random 128-byte line per lane, every lane in bank group 0 (worst case)
//   4  lanes 4k..4k+3 read the four records of one random line (8 lines per instruction)
__device__ __forceinline__ u32 mix32(u32 x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
template<int PATTERN>
__global__ void __launch_bounds__(128) gather32_kernel(const Pack4 *__restrict__ tab, u32 mask, int iters, f64 *out) {
    const u32 lane = threadIdx.x & 31u;
    u32 state      = (blockIdx.x * blockDim.x + threadIdx.x) / 32u * 2654435761u + 12345u;
    f64 acc        = 0;
    for (int i = 0; i < iters; i++) {
        state = state * 1664525u + 1013904223u;
        u32 idx;
        if (PATTERN == 0)
            idx = (state & ~31u) + lane;
        else if (PATTERN == 1)
            idx = mix32(state ^ (lane * 0x9E3779B9u));
        else if (PATTERN == 2)
            idx = (mix32(state ^ (lane * 0x9E3779B9u)) & ~3u) | (lane & 3u);
        else if (PATTERN == 3)
            idx = mix32(state ^ (lane * 0x9E3779B9u)) & ~3u;
        else
            idx = (mix32(state ^ ((lane >> 2) * 0x9E3779B9u)) & ~3u) | (lane & 3u);
        const Pack4 r = ld4(tab + (idx & mask));
        acc += r.a + r.d;
    }
    if (acc == 12345.678) // never true: keeps the loads alive
        out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

/// G records / s of the whole GPU for one pattern; log2_records = table size (10: 32 KB, L1 resident)
static f64 gather_bench(Ctx &c, int pattern, int log2_records, cudaEvent_t e0, cudaEvent_t e1) {
    cudaStream_t s   = c.stream;
    const u64 nrec   = u64(1) << log2_records;
    const int iters  = 2048;
    const int blocks = kNumSM * 8;
    DevBuf<Pack4> tab;
    DevBuf<f64> out;
    tab.ensure(nrec);
    out.ensure(size_t(blocks) * 128);
    SB_CUDA_CHECK(cudaMemsetAsync(tab.p, 0, nrec * sizeof(Pack4), s));
    f64 best = 0;
    for (int rep = 0; rep < 4; rep++) {
        SB_CUDA_CHECK(cudaEventRecord(e0, s));
        switch (pattern) {
        case 0: gather32_kernel<0><<<blocks, 128, 0, s>>>(tab.p, u32(nrec - 1), iters, out.p); break;
        case 1: gather32_kernel<1><<<blocks, 128, 0, s>>>(tab.p, u32(nrec - 1), iters, out.p); break;
        case 2: gather32_kernel<2><<<blocks, 128, 0, s>>>(tab.p, u32(nrec - 1), iters, out.p); break;
        case 3: gather32_kernel<3><<<blocks, 128, 0, s>>>(tab.p, u32(nrec - 1), iters, out.p); break;
        default: gather32_kernel<4><<<blocks, 128, 0, s>>>(tab.p, u32(nrec - 1), iters, out.p); break;
        }
        SB_CUDA_CHECK(cudaEventRecord(e1, s));
        SB_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0)
            best = std::max(best, f64(blocks) * 128 * f64(iters) / (ms * 1e-3) / 1e9);
    }
    SB_LAUNCH_CHECK();
    return best;
}

/// what: 0 = FP64 FMA TFLOP/s (FMA = 2 flop), 1 = copy GB/s (read + write bytes),
/// 10 + p = 32-byte record gathers, pattern p, L1-resident table [G records / s],
/// 20 + p = the same from a 32 MiB (L2-resident) table
f64 microbench(Ctx &c, int what) {
    cudaStream_t s = c.stream;
    cudaEvent_t e0, e1;
    SB_CUDA_CHECK(cudaEventCreate(&e0));
    SB_CUDA_CHECK(cudaEventCreate(&e1));
    f64 best = 0;
    if (what >= 10 && what < 30) {
        best = gather_bench(c, what % 10, what >= 20 ? 20 : 10, e0, e1);
    } else if (what == 0) {
        constexpr int CH = 8;
        const int iters  = 4096;
        const int blocks = kNumSM * 16;
        DevBuf<f64> out;
        out.ensure(size_t(blocks) * 256);
        for (int rep = 0; rep < 5; rep++) {
            SB_CUDA_CHECK(cudaEventRecord(e0, s));
            fma_chain_kernel<CH><<<blocks, 256, 0, s>>>(out.p, iters, 1.0000001, 1e-9);
            SB_CUDA_CHECK(cudaEventRecord(e1, s));
            SB_CUDA_CHECK(cudaEventSynchronize(e1));
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            f64 flops = f64(blocks) * 256 * CH * f64(iters) * 2;
            if (rep > 0)
                best = std::max(best, flops / (ms * 1e-3) / 1e12);
        }
    } else {
        const u64 n = (u64(1) << 30) / sizeof(double2); // 1 GiB each way
        DevBuf<double2> a, b;
        a.ensure(n);
        b.ensure(n);
        SB_CUDA_CHECK(cudaMemsetAsync(a.p, 0, n * sizeof(double2), s));
        for (int rep = 0; rep < 6; rep++) {
            SB_CUDA_CHECK(cudaEventRecord(e0, s));
            stream_copy_kernel<<<kNumSM * 16, 256, 0, s>>>(a.p, b.p, n);
            SB_CUDA_CHECK(cudaEventRecord(e1, s));
            SB_CUDA_CHECK(cudaEventSynchronize(e1));
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0)
                best = std::max(best, 2.0 * f64(n) * sizeof(double2) / (ms * 1e-3) / 1e9);
        }
    }
    SB_COUNT_LAUNCH();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

} // namespace sb
