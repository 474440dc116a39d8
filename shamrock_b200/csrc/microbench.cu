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

/// what: 0 = FP64 FMA TFLOP/s (FMA = 2 flop), 1 = copy GB/s (read + write bytes)
f64 microbench(Ctx &c, int what) {
    cudaStream_t s = c.stream;
    cudaEvent_t e0, e1;
    SB_CUDA_CHECK(cudaEventCreate(&e0));
    SB_CUDA_CHECK(cudaEventCreate(&e1));
    f64 best = 0;
    if (what == 0) {
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
