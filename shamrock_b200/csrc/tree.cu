// tree.cu — Morton codes, key/value sort (bitonic-exact and radix), leaf compression, Karras radix
// tree, AABB / max-field propagation.  Hand-written CUDA for sm_100a, no CUB.
//
// Compiled with -fmad=false: the Morton transform and every float comparison here must round like
// the reference's separate IEEE operations (bit-exact contract).
//
// Reference behaviour restated here (paths relative to /root/reference/src):
//   shamtree/src/MortonCodeSet.cpp:61-128, shammath/include/shammath/sfc/{morton,bmi}.hpp,
//   shammath/src/CoordRangeTransform.cpp:169-184, shamalgs/src/details/algorithm/
//   bitonicSort_updated_usm.cpp:29-39,285-396, shamtree/src/kernels/reduction_alg.cpp:63-85,
//   197-244,274-295,565-585, shamtree/src/KarrasRadixTree.cpp:46-143,
//   shamtree/src/KarrasRadixTreeAABB.cpp:33-135, shamtree/include/shamtree/KarrasRadixTreeField.hpp:134-222,
//   shammodels/sph/src/modules/BuildTrees.cpp:37-56.
#include "tree.cuh"
#include <cstring>

namespace sb {

// =============================================================================================
// bounding box (K1) : min/max of the positions, widened by one ulp (BuildTrees.cpp:41-56)
// =============================================================================================
__global__ void bbox_init_kernel(u64 *acc) {
    if (threadIdx.x < 3)
        acc[threadIdx.x] = 0xFFFFFFFFFFFFFFFFull; // min accumulators
    else if (threadIdx.x < 6)
        acc[threadIdx.x] = 0ull; // max accumulators
}

__global__ void __launch_bounds__(256) bbox_reduce_kernel(
    const f64 *__restrict__ xyz, size_t stride, u32 n, u64 *acc) {
    f64 mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += u64(gridDim.x) * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            f64 v = xyz[i * stride + c];
            mn[c] = fmin(mn[c], v);
            mx[c] = fmax(mx[c], v);
        }
    }
    __shared__ f64 smn[8][3], smx[8][3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        mn[c] = warp_min(mn[c]);
        mx[c] = warp_max(mx[c]);
    }
    int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            smn[w][c] = mn[c];
            smx[w][c] = mx[c];
        }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        int c  = threadIdx.x;
        f64 a = smn[0][c], b = smx[0][c];
        for (int k = 1; k < 8; k++) {
            a = fmin(a, smn[k][c]);
            b = fmax(b, smx[k][c]);
        }
        atomicMin((unsigned long long *) &acc[c], (unsigned long long) f64_to_ordered(a));
        atomicMax((unsigned long long *) &acc[3 + c], (unsigned long long) f64_to_ordered(b));
    }
}

/// widen == true: nextafter(min, -inf), nextafter(max, +inf)
__global__ void bbox_finalize_kernel(const u64 *acc, f64 *bbox, bool widen) {
    int c = threadIdx.x;
    if (c < 3) {
        f64 a   = ordered_to_f64(acc[c]);
        f64 b   = ordered_to_f64(acc[3 + c]);
        bbox[c] = widen ? ::nextafter(a, -f64(INFINITY)) : a;
        bbox[3 + c] = widen ? ::nextafter(b, f64(INFINITY)) : b;
    }
}

// =============================================================================================
// Morton codes (K2-K4)
// =============================================================================================
__device__ __forceinline__ u32 expand_bits_10(u32 x) {
    x &= 0x3ffU;
    x = (x | x << 16U) & 0x30000ffU;
    x = (x | x << 8U) & 0x300f00fU;
    x = (x | x << 4U) & 0x30c30c3U;
    x = (x | x << 2U) & 0x9249249U;
    return x;
}

__global__ void __launch_bounds__(256) morton_kernel(
    const f64 *__restrict__ xyz, size_t stride, u32 cnt_obj, u32 morton_count,
    const f64 *__restrict__ bbox, u32 *__restrict__ morton, u32 *__restrict__ index_map) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= morton_count)
        return;
    index_map[i] = i;
    if (i >= cnt_obj) {
        morton[i] = 0xFFFFFFFFu;
        return;
    }
    u32 ic[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        f64 lo = bbox[c], hi = bbox[3 + c];
        f64 fact = (hi - lo) / 1024.0;
        f64 r    = fmin(fmax(xyz[u64(i) * stride + c], lo), hi); // sycl::clamp
        // static_cast<u16>((r - lo) / fact) then clamp to [0, 1023]
        u32 v = (u32) (unsigned short) __double2int_rz((r - lo) / fact);
        ic[c] = min(max(v, 0u), 1023u);
    }
    morton[i] = expand_bits_10(ic[0]) * 4 + expand_bits_10(ic[1]) * 2 + expand_bits_10(ic[2]);
}

// =============================================================================================
// Bitonic network (K5) : identical compare-exchange semantics as the reference
//   swap = reverse ^ (a < b), reverse = ((2*length) & i) == 0
// =============================================================================================
__device__ __forceinline__ void cmpxchg(u32 &a, u32 &b, u32 &va, u32 &vb, bool reverse) {
    bool swap = reverse ^ (a < b);
    u32 ta = a, tb = b, tva = va, tvb = vb;
    a  = swap ? tb : ta;
    b  = swap ? ta : tb;
    va = swap ? tvb : tva;
    vb = swap ? tva : tvb;
}

constexpr int BITONIC_LOG_TILE = 12; // 4096 (key,value) pairs = 32 KB of shared memory per CTA
constexpr int BITONIC_TILE     = 1 << BITONIC_LOG_TILE;
constexpr int BITONIC_THREADS  = 1024;

/// Shared-memory stages.  full_sort: all (length, inc) with length < tile_len; otherwise only the
/// tail inc = tile_len/2 .. 1 of the given `length`.
__global__ void __launch_bounds__(BITONIC_THREADS) bitonic_smem_kernel(
    u32 *__restrict__ keys, u32 *__restrict__ vals, u32 tile_len, u32 length_arg, bool full_sort) {
    __shared__ u32 sk[BITONIC_TILE];
    __shared__ u32 sv[BITONIC_TILE];
    const u32 base = blockIdx.x * tile_len;
    for (u32 j = threadIdx.x; j < tile_len; j += BITONIC_THREADS) {
        sk[j] = keys[base + j];
        sv[j] = vals[base + j];
    }
    __syncthreads();
    u32 length_begin = full_sort ? 1u : length_arg;
    u32 length_end   = full_sort ? tile_len : (length_arg << 1); // exclusive
    for (u32 length = length_begin; length < length_end; length <<= 1) {
        u32 inc0 = full_sort ? length : (tile_len >> 1);
        u32 dir  = length << 1;
        for (u32 inc = inc0; inc > 0; inc >>= 1) {
            for (u32 t = threadIdx.x; t < (tile_len >> 1); t += BITONIC_THREADS) {
                u32 low = t & (inc - 1);
                u32 i   = (t << 1) - low;
                bool reverse = ((dir & (base + i)) == 0);
                u32 a = sk[i], b = sk[i + inc], va = sv[i], vb = sv[i + inc];
                cmpxchg(a, b, va, vb, reverse);
                sk[i]       = a;
                sk[i + inc] = b;
                sv[i]       = va;
                sv[i + inc] = vb;
            }
            __syncthreads();
        }
    }
    for (u32 j = threadIdx.x; j < tile_len; j += BITONIC_THREADS) {
        keys[base + j] = sk[j];
        vals[base + j] = sv[j];
    }
}

/// Global stages: F fused levels (inc, inc/2, .. inc >> (F-1)), 2^F elements per thread held in
/// registers (the reference's order_kernel<2^F> stencils).
template<int F>
__global__ void __launch_bounds__(256) bitonic_global_kernel(
    u32 *__restrict__ keys, u32 *__restrict__ vals, u32 inc, u32 length, u32 nthreads) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads)
        return;
    constexpr int E = 1 << F;
    u32 _inc        = inc >> (F - 1);
    u32 low         = t & (_inc - 1);
    u32 i           = ((t - low) << F) + low;
    bool reverse    = (((length << 1) & i) == 0);
    u32 x[E], v[E];
#pragma unroll
    for (int k = 0; k < E; k++) {
        x[k] = keys[i + k * _inc];
        v[k] = vals[i + k * _inc];
    }
#pragma unroll
    for (int s = E >> 1; s > 0; s >>= 1) {
#pragma unroll
        for (int k = 0; k < E; k++) {
            if ((k & s) == 0)
                cmpxchg(x[k], x[k + s], v[k], v[k + s], reverse);
        }
    }
#pragma unroll
    for (int k = 0; k < E; k++) {
        keys[i + k * _inc] = x[k];
        vals[i + k * _inc] = v[k];
    }
}

void bitonic_sort_by_key(cudaStream_t s, u32 *keys, u32 *vals, u32 len) {
    if (len & (len - 1))
        throw std::invalid_argument("this algorithm can only be used with length that are powers of two");
    if (len <= 1)
        return;
    u32 tile = len < (u32) BITONIC_TILE ? len : (u32) BITONIC_TILE;
    u32 nb   = len / tile;
    bitonic_smem_kernel<<<nb, BITONIC_THREADS, 0, s>>>(keys, vals, tile, 0, true);
    SB_COUNT_LAUNCH();
    for (u32 length = tile; length < len; length <<= 1) {
        u32 inc = length;
        // global levels: inc = length .. tile (those with inc >= tile)
        while (inc >= tile) {
            int levels_left = 0;
            for (u32 q = inc; q >= tile; q >>= 1)
                levels_left++;
            if (levels_left >= 3) {
                u32 nt = len >> 3;
                bitonic_global_kernel<3><<<grid_for(nt, 256), 256, 0, s>>>(keys, vals, inc, length, nt);
                inc >>= 3;
            } else if (levels_left == 2) {
                u32 nt = len >> 2;
                bitonic_global_kernel<2><<<grid_for(nt, 256), 256, 0, s>>>(keys, vals, inc, length, nt);
                inc >>= 2;
            } else {
                u32 nt = len >> 1;
                bitonic_global_kernel<1><<<grid_for(nt, 256), 256, 0, s>>>(keys, vals, inc, length, nt);
                inc >>= 1;
            }
            SB_COUNT_LAUNCH();
        }
        bitonic_smem_kernel<<<nb, BITONIC_THREADS, 0, s>>>(keys, vals, tile, length, false);
        SB_COUNT_LAUNCH();
    }
    SB_LAUNCH_CHECK();
}

// =============================================================================================
// LSD radix sort, 8-bit digits, stable (CUB-free).  Same sorted keys as the network; ties keep the
// input order (≠ the network's tie order) — used when only the tree topology
// matters (tree micro-bench, perf mode).
// =============================================================================================
constexpr int RADIX_THREADS = 256;
constexpr int RADIX_ITEMS   = 16;
constexpr int RADIX_TILE    = RADIX_THREADS * RADIX_ITEMS; // 4096 keys per CTA, 512 per warp

__global__ void __launch_bounds__(RADIX_THREADS) radix_hist_kernel(
    const u32 *__restrict__ keys, u32 n, int shift, u32 nblocks, u32 *__restrict__ hist) {
    __shared__ u32 sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    u32 base = blockIdx.x * RADIX_TILE;
#pragma unroll
    for (int k = 0; k < RADIX_ITEMS; k++) {
        u32 i = base + k * RADIX_THREADS + threadIdx.x;
        if (i < n)
            atomicAdd(&sh[(keys[i] >> shift) & 0xFF], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * nblocks + blockIdx.x] = sh[threadIdx.x];
}

__global__ void __launch_bounds__(RADIX_THREADS) radix_scatter_kernel(
    const u32 *__restrict__ keys, const u32 *__restrict__ vals, u32 n, int shift, u32 nblocks,
    const u32 *__restrict__ offsets, u32 *__restrict__ keys_out, u32 *__restrict__ vals_out) {
    constexpr int NW = RADIX_THREADS / 32;
    __shared__ u32 whist[NW][256];
    for (int j = threadIdx.x; j < NW * 256; j += RADIX_THREADS)
        (&whist[0][0])[j] = 0;
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 wbase = blockIdx.x * RADIX_TILE + w * (RADIX_TILE / NW);
    u32 k[RADIX_ITEMS], v[RADIX_ITEMS];
    u32 rank[RADIX_ITEMS];
#pragma unroll
    for (int r = 0; r < RADIX_ITEMS; r++) {
        u32 i      = wbase + r * 32 + lane;
        bool valid = i < n;
        k[r]       = valid ? keys[i] : 0xFFFFFFFFu;
        v[r]       = valid ? vals[i] : 0u;
        u32 d      = (k[r] >> shift) & 0xFF;
        // invalid lanes get a digit outside the table so that they never match valid ones
        u32 md      = valid ? d : 256u + lane;
        u32 mask    = __match_any_sync(0xffffffffu, md);
        u32 before  = __popc(mask & ((1u << lane) - 1u));
        int leader  = __ffs(mask) - 1;
        u32 pre     = 0;
        if (valid && lane == leader) {
            pre         = whist[w][d];
            whist[w][d] = pre + __popc(mask);
        }
        pre     = __shfl_sync(0xffffffffu, pre, leader);
        rank[r] = pre + before;
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive prefix over the warps of this CTA; then the CTA-local start of every digit
    // (block scan over the 256 digit counts) and the global offset of the CTA's run of that digit
    __shared__ u32 dstart[256], gofs[256], wtot[NW];
    __shared__ u32 skey[RADIX_TILE], sval[RADIX_TILE];
    {
        u32 d   = threadIdx.x;
        u32 run = 0;
#pragma unroll
        for (int ww = 0; ww < NW; ww++) {
            u32 c        = whist[ww][d];
            whist[ww][d] = run;
            run += c;
        }
        u32 inc = run; // inclusive scan of the digit counts inside the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o)
                inc += t;
        }
        if (lane == 31)
            wtot[w] = inc;
        __syncthreads();
        u32 before = 0;
#pragma unroll
        for (int ww = 0; ww < NW; ww++)
            before += ww < w ? wtot[ww] : 0u;
        dstart[d] = before + inc - run;
        gofs[d]   = offsets[d * nblocks + blockIdx.x] - (before + inc - run);
    }
    __syncthreads();
    // stage the tile in digit order in shared memory (stable: warp, round, lane = input order) ...
#pragma unroll
    for (int r = 0; r < RADIX_ITEMS; r++) {
        u32 i = wbase + r * 32 + lane;
        if (i < n) {
            u32 d    = (k[r] >> shift) & 0xFF;
            u32 lp   = dstart[d] + whist[w][d] + rank[r];
            skey[lp] = k[r];
            sval[lp] = v[r];
        }
    }
    __syncthreads();
    // ... and write it out: consecutive threads write consecutive addresses inside each digit's run
    const u32 tile0 = blockIdx.x * RADIX_TILE;
    const u32 ntile = n - tile0 < u32(RADIX_TILE) ? n - tile0 : u32(RADIX_TILE);
    for (u32 i = threadIdx.x; i < ntile; i += RADIX_THREADS) {
        u32 kk  = skey[i];
        u32 pos = gofs[(kk >> shift) & 0xFF] + i;
        keys_out[pos] = kk;
        vals_out[pos] = sval[i];
    }
}

void radix_sort_by_key(
    cudaStream_t s, u32 *keys, u32 *vals, u32 *keys_alt, u32 *vals_alt, u32 len, int bits,
    DevBuf<u32> &hist) {
    if (len <= 1)
        return;
    u32 nb = (len + RADIX_TILE - 1) / RADIX_TILE;
    hist.ensure(size_t(256) * nb * 2 + 4096);
    u32 *h_in  = hist.p;
    u32 *h_out = hist.p + size_t(256) * nb;
    DevBuf<u32> scan_tmp;
    DevBuf<u64> tot;
    tot.ensure(1);
    u32 *ki = keys, *vi = vals, *ko = keys_alt, *vo = vals_alt;
    int passes = (bits + 7) / 8;
    if (passes & 1)
        passes++; // even number of passes so that the result lands in keys/vals
    for (int p = 0; p < passes; p++) {
        int shift = 8 * p;
        radix_hist_kernel<<<nb, RADIX_THREADS, 0, s>>>(ki, len, shift, nb, h_in);
        SB_COUNT_LAUNCH();
        exclusive_scan<u32>(s, h_in, h_out, u64(256) * nb, scan_tmp, tot.p);
        radix_scatter_kernel<<<nb, RADIX_THREADS, 0, s>>>(ki, vi, len, shift, nb, h_out, ko, vo);
        SB_COUNT_LAUNCH();
        std::swap(ki, ko);
        std::swap(vi, vo);
    }
    SB_LAUNCH_CHECK();
    SB_CUDA_CHECK(cudaStreamSynchronize(s)); // scan_tmp / tot are freed at scope exit
}

// =============================================================================================
// Onesweep LSD radix sort (CUB-free): ONE histogram kernel reads the keys once for all digits, then one kernel
// per 8-bit digit ranks a tile of 4096 keys, publishes the tile's digit counts and finds its global offsets by
// decoupled look-back over the status words of the tiles before it (Adinets & Merrill 2022; the reference's own
// experimental version: shamalgs/include/shamalgs/details/algorithm/radixSortOnesweep.hpp:53) — no per-pass
// histogram / scan kernels, no host synchronisation.  Stable; same result as radix_sort_by_key.
// Status word: [flag:2 | value:30], flag 1 = the tile's own count (aggregate), 2 = inclusive prefix.
// =============================================================================================
constexpr u32 OS_FLAG_AGG = 1u << 30, OS_FLAG_PRE = 2u << 30, OS_VALUE = (1u << 30) - 1u;

__device__ __forceinline__ u32 ld_relaxed_gpu(const u32 *p) {
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(u32 *p, u32 v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

/// all digit histograms in one pass over the keys: hist[pass * 256 + digit]
__global__ void __launch_bounds__(256) onesweep_hist_kernel(const u32 *__restrict__ keys, u32 n, int passes, u32 *__restrict__ hist) {
    __shared__ u32 sh[4 * 256];
    for (int j = threadIdx.x; j < 4 * 256; j += 256)
        sh[j] = 0;
    __syncthreads();
    for (u64 i = u64(blockIdx.x) * 256 + threadIdx.x; i < n; i += u64(gridDim.x) * 256) {
        const u32 k = keys[i];
        for (int p = 0; p < passes; p++)
            atomicAdd(&sh[p * 256 + ((k >> (8 * p)) & 0xFF)], 1u);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < passes * 256; j += 256)
        if (sh[j])
            atomicAdd(&hist[j], sh[j]);
}
/// exclusive scan of every pass's 256 digit totals (one warp-scan block per pass), in place
__global__ void __launch_bounds__(256) onesweep_bases_kernel(u32 *__restrict__ hist) {
    __shared__ u32 wsum[8];
    u32 *h   = hist + blockIdx.x * 256;
    u32 v    = h[threadIdx.x];
    u32 inc  = v;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o)
            inc += t;
    }
    if (lane == 31)
        wsum[w] = inc;
    __syncthreads();
    u32 before = 0;
#pragma unroll
    for (int ww = 0; ww < 8; ww++)
        before += ww < w ? wsum[ww] : 0u;
    h[threadIdx.x] = before + inc - v;
}

template<bool IOTA>
__global__ void __launch_bounds__(RADIX_THREADS) onesweep_pass_kernel(
    const u32 *__restrict__ keys, const u32 *__restrict__ vals, u32 n, int shift, const u32 *__restrict__ gbase,
    u32 *__restrict__ status, u32 *__restrict__ tile_counter, u32 *__restrict__ keys_out, u32 *__restrict__ vals_out) {
    constexpr int NW = RADIX_THREADS / 32;
    __shared__ u32 whist[NW][256];
    __shared__ u32 dstart[256], gofs[256], wtot[NW];
    __shared__ u32 skey[RADIX_TILE], sval[RADIX_TILE];
    __shared__ u32 s_tile;
    if (threadIdx.x == 0)
        s_tile = atomicAdd(tile_counter, 1u); // tiles are numbered in the order they start: a tile only waits
    for (int j = threadIdx.x; j < NW * 256; j += RADIX_THREADS) // for tiles that are already running
        (&whist[0][0])[j] = 0;
    __syncthreads();
    const u32 tile = s_tile;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 wbase = tile * RADIX_TILE + w * (RADIX_TILE / NW);
    u32 k[RADIX_ITEMS], v[RADIX_ITEMS];
    u32 rank[RADIX_ITEMS];
#pragma unroll
    for (int r = 0; r < RADIX_ITEMS; r++) {
        u32 i      = wbase + r * 32 + lane;
        bool valid = i < n;
        k[r]       = valid ? keys[i] : 0xFFFFFFFFu;
        v[r]       = valid ? (IOTA ? i : vals[i]) : 0u;
        u32 d      = (k[r] >> shift) & 0xFF;
        u32 md      = valid ? d : 256u + lane; // invalid lanes never match valid ones
        u32 mask    = __match_any_sync(0xffffffffu, md);
        u32 before  = __popc(mask & ((1u << lane) - 1u));
        int leader  = __ffs(mask) - 1;
        u32 pre     = 0;
        if (valid && lane == leader) {
            pre         = whist[w][d];
            whist[w][d] = pre + __popc(mask);
        }
        pre     = __shfl_sync(0xffffffffu, pre, leader);
        rank[r] = pre + before;
        __syncwarp();
    }
    __syncthreads();
    {
        const u32 d = threadIdx.x;
        u32 run     = 0;
#pragma unroll
        for (int ww = 0; ww < NW; ww++) {
            u32 c        = whist[ww][d];
            whist[ww][d] = run;
            run += c;
        }
        // publish this tile's count of digit d, then sum the tiles before it (decoupled look-back)
        u32 *st = status + size_t(tile) * 256 + d;
        st_relaxed_gpu(st, (tile == 0 ? OS_FLAG_PRE : OS_FLAG_AGG) | run);
        u32 excl = 0;
        if (tile > 0) {
            for (u32 t = tile - 1;;) {
                const u32 sv = ld_relaxed_gpu(status + size_t(t) * 256 + d);
                if ((sv >> 30) == 0)
                    continue; // not published yet: that tile is running, spin
                excl += sv & OS_VALUE;
                if (sv & OS_FLAG_PRE)
                    break;
                t--; // an aggregate: keep walking (tile 0 always publishes a prefix)
            }
            st_relaxed_gpu(st, OS_FLAG_PRE | (excl + run));
        }
        u32 inc = run; // tile-local start of every digit: block scan over the 256 digit counts
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o)
                inc += t;
        }
        if (lane == 31)
            wtot[w] = inc;
        __syncthreads();
        u32 before = 0;
#pragma unroll
        for (int ww = 0; ww < NW; ww++)
            before += ww < w ? wtot[ww] : 0u;
        dstart[d] = before + inc - run;
        gofs[d]   = gbase[d] + excl - (before + inc - run);
    }
    __syncthreads();
    // stage the tile in digit order in shared memory (stable: warp, round, lane = input order) ...
#pragma unroll
    for (int r = 0; r < RADIX_ITEMS; r++) {
        u32 i = wbase + r * 32 + lane;
        if (i < n) {
            u32 d    = (k[r] >> shift) & 0xFF;
            u32 lp   = dstart[d] + whist[w][d] + rank[r];
            skey[lp] = k[r];
            sval[lp] = v[r];
        }
    }
    __syncthreads();
    // ... and write it out: consecutive threads write consecutive addresses inside each digit's run
    const u32 tile0 = tile * RADIX_TILE;
    const u32 ntile = n - tile0 < u32(RADIX_TILE) ? n - tile0 : u32(RADIX_TILE);
    for (u32 i = threadIdx.x; i < ntile; i += RADIX_THREADS) {
        u32 kk  = skey[i];
        u32 pos = gofs[(kk >> shift) & 0xFF] + i;
        keys_out[pos] = kk;
        vals_out[pos] = sval[i];
    }
}

/// stable onesweep sort on `bits` key bits; values = 0..len-1 when iota_values (the value array is then not
/// read in the first pass).  Result in keys / vals.  No host synchronisation.
void onesweep_sort_by_key(
    cudaStream_t s, u32 *keys, u32 *vals, u32 *keys_alt, u32 *vals_alt, u32 len, int bits, DevBuf<u32> &work,
    bool iota_values) {
    if (len <= 1)
        return;
    if (len >= (1u << 30)) { // status words carry 30-bit prefixes
        radix_sort_by_key(s, keys, vals, keys_alt, vals_alt, len, bits, work);
        return;
    }
    int passes = (bits + 7) / 8;
    if (passes & 1)
        passes++; // even number of passes so that the result lands in keys / vals
    if (passes > 4)
        throw std::invalid_argument("onesweep: at most 32 key bits");
    const u32 nt = (len + RADIX_TILE - 1) / RADIX_TILE;
    // work: [4 * 256 histograms / bases | 4 tile counters (+ pad) | passes * nt * 256 status words]
    const size_t words = 4 * 256 + 16 + size_t(passes) * nt * 256;
    work.ensure(words);
    SB_CUDA_CHECK(cudaMemsetAsync(work.p, 0, words * sizeof(u32), s));
    u32 *hist = work.p, *counters = work.p + 4 * 256, *status = work.p + 4 * 256 + 16;
    unsigned nb = (unsigned) std::min<u64>(u64(kNumSM) * 8, (u64(len) + 255) / 256);
    onesweep_hist_kernel<<<nb, 256, 0, s>>>(keys, len, passes, hist);
    SB_COUNT_LAUNCH();
    onesweep_bases_kernel<<<passes, 256, 0, s>>>(hist);
    SB_COUNT_LAUNCH();
    u32 *ki = keys, *vi = vals, *ko = keys_alt, *vo = vals_alt;
    for (int p = 0; p < passes; p++) {
        if (p == 0 && iota_values)
            onesweep_pass_kernel<true><<<nt, RADIX_THREADS, 0, s>>>(
                ki, vi, len, 8 * p, hist + 256 * p, status + size_t(p) * nt * 256, counters + p, ko, vo);
        else
            onesweep_pass_kernel<false><<<nt, RADIX_THREADS, 0, s>>>(
                ki, vi, len, 8 * p, hist + 256 * p, status + size_t(p) * nt * 256, counters + p, ko, vo);
        SB_COUNT_LAUNCH();
        std::swap(ki, ko);
        std::swap(vi, vo);
    }
    SB_LAUNCH_CHECK();
}

// =============================================================================================
// Leaf compression (K6-K9)
// =============================================================================================
__device__ __forceinline__ int karras_delta_d(int x, int y, u32 morton_length, const u32 *m) {
    return ((u32(y) > morton_length - 1 || y < 0) ? -1 : __clz(int(m[x] ^ m[y])));
}

__global__ void __launch_bounds__(256) split_table_kernel(const u32 *__restrict__ m, u32 n, u8 *__restrict__ split) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    split[i] = (i > 0) ? ((m[i - 1] != m[i]) ? 1 : 0) : 1;
}

__global__ void __launch_bounds__(256) reduction_iteration_kernel(
    const u32 *__restrict__ m, u32 n, const u8 *__restrict__ split_in, u8 *__restrict__ split_out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    u32 before1 = i - 1;
    while (before1 <= n - 1 && !split_in[before1])
        before1--;
    u32 before2 = before1 - 1;
    while (before2 <= n - 1 && !split_in[before2])
        before2--;
    u32 next1 = i + 1;
    while (next1 <= n - 1 && !split_in[next1])
        next1++;
    int delt_0  = karras_delta_d(int(i), int(next1), n, m);
    int delt_m  = karras_delta_d(int(i), int(before1), n, m);
    int delt_mm = karras_delta_d(int(before1), int(before2), n, m);
    split_out[i] = (!(delt_0 < delt_m && delt_mm < delt_m) && split_in[i]) ? 1 : 0;
}

__global__ void __launch_bounds__(256) compact_leaves_kernel(
    const u32 *__restrict__ m, u32 n, const u8 *__restrict__ split, const u32 *__restrict__ pos,
    const u64 *__restrict__ d_total, u32 *__restrict__ reduc_index_map, u32 *__restrict__ reduced_morton) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        u32 L                  = u32(*d_total);
        reduc_index_map[L]     = n;
        reduc_index_map[L + 1] = 0;
    }
    if (i >= n)
        return;
    if (split[i]) {
        u32 p              = pos[i];
        reduc_index_map[p] = i;
        reduced_morton[p]  = m[i];
    }
}

// =============================================================================================
// Karras 2012 (K10), + parent pointers for the bottom-up passes
// =============================================================================================
__global__ void __launch_bounds__(256) karras_kernel(
    const u32 *__restrict__ morton, u32 internal_cell_count, u32 *__restrict__ lchild_id,
    u32 *__restrict__ rchild_id, u8 *__restrict__ lchild_flag, u8 *__restrict__ rchild_flag,
    u32 *__restrict__ end_range_cell, u32 *__restrict__ parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= int(internal_cell_count))
        return;
    const u32 morton_length = internal_cell_count + 1;
    auto DELTA = [&](int x, int y) { return karras_delta_d(x, y, morton_length, morton); };
    int ddelta = DELTA(i, i + 1) - DELTA(i, i - 1);
    int d      = (ddelta == 0) ? 0 : ((ddelta > 0) ? 1 : -1);
    int delta_min = DELTA(i, i - d);
    int lmax      = 2;
    while (DELTA(i, i + lmax * d) > delta_min)
        lmax *= 2;
    int l = 0;
    int t = lmax / 2;
    while (t > 0) {
        if (DELTA(i, i + (l + t) * d) > delta_min)
            l = l + t;
        t = t / 2;
    }
    u32 j             = u32(i + l * d);
    end_range_cell[i] = j;
    int delta_node    = DELTA(i, int(j));
    int s             = 0;
    float div         = 2;
    t                 = int(ceilf(float(l) / div)); // the reference's `float div` quirk, kept
    while (true) {
        int tmp_ = i + (s + t) * d;
        if (DELTA(i, tmp_) > delta_node)
            s = s + t;
        if (t <= 1)
            break;
        div *= 2;
        t = int(ceilf(float(l) / div));
    }
    int gamma = i + s * d + min(d, 0);
    u8 lf     = (min(i, int(j)) == gamma) ? 1 : 0;
    u8 rf     = (max(i, int(j)) == gamma + 1) ? 1 : 0;
    lchild_id[i]   = u32(gamma);
    lchild_flag[i] = lf;
    rchild_id[i]   = u32(gamma + 1);
    rchild_flag[i] = rf;
    parent[u32(gamma) + internal_cell_count * lf]     = u32(i);
    parent[u32(gamma + 1) + internal_cell_count * rf] = u32(i);
}

// =============================================================================================
// AABB (K11-K12) and max-field (K13): one bottom-up pass with arrival counters.  min/max are exact,
// so this equals the reference's 32 brute-force passes (KarrasRadixTreeAABB.cpp:33-74).
// =============================================================================================
/// FIELD: the objects are (x, y, z, f) records and the pass also reduces fout[node] = fscale * max f, exactly as
/// leaf_field_max_propagate_kernel would afterwards (the solver's interaction radius: compute_presteps_rint)
template<bool FIELD>
__global__ void __launch_bounds__(128) leaf_aabb_propagate_kernel(
    const f64 *__restrict__ xyz, size_t stride, const u32 *__restrict__ index_map,
    const u32 *__restrict__ reduc_index_map, u32 L, u32 I, const u32 *__restrict__ lchild,
    const u32 *__restrict__ rchild, const u8 *__restrict__ lflag, const u8 *__restrict__ rflag,
    const u32 *__restrict__ parent, u32 *__restrict__ counters, f64 *aabb_min, f64 *aabb_max, f64 fscale,
    f64 *fout) {
    u32 leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= L)
        return;
    const f64 big = 1.7976931348623157e308;
    f64 mn[3] = {big, big, big}, mx[3] = {-big, -big, -big};
    f64 fm = -big;
    u32 a = reduc_index_map[leaf], b = reduc_index_map[leaf + 1];
    for (u32 sidx = a; sidx < b; sidx++) {
        u32 id = index_map[sidx];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            f64 v = xyz[u64(id) * stride + c];
            mn[c] = fmin(mn[c], v);
            mx[c] = fmax(mx[c], v);
        }
        if (FIELD)
            fm = fmax(fm, xyz[u64(id) * stride + 3]);
    }
    u32 node = I + leaf;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        aabb_min[u64(node) * 3 + c] = mn[c];
        aabb_max[u64(node) * 3 + c] = mx[c];
    }
    if (FIELD)
        fout[node] = fm * fscale;
    if (I == 0)
        return;
    node = parent[node];
    while (true) {
        __threadfence();
        u32 old = atomicAdd(&counters[node], 1u);
        if (old == 0)
            return; // the sibling subtree is not finished: its last thread will continue
        u32 l = lchild[node] + I * u32(lflag[node]);
        u32 r = rchild[node] + I * u32(rflag[node]);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            f64 a0 = __ldcg(&aabb_min[u64(l) * 3 + c]), a1 = __ldcg(&aabb_min[u64(r) * 3 + c]);
            f64 b0 = __ldcg(&aabb_max[u64(l) * 3 + c]), b1 = __ldcg(&aabb_max[u64(r) * 3 + c]);
            aabb_min[u64(node) * 3 + c] = fmin(a0, a1);
            aabb_max[u64(node) * 3 + c] = fmax(b0, b1);
        }
        if (FIELD)
            fout[node] = fmax(__ldcg(&fout[l]), __ldcg(&fout[r]));
        if (node == 0)
            return;
        node = parent[node];
    }
}

__global__ void __launch_bounds__(128) leaf_field_max_propagate_kernel(
    const f64 *__restrict__ field, size_t fstride, const u32 *__restrict__ index_map,
    const u32 *__restrict__ reduc_index_map, u32 L, u32 I, const u32 *__restrict__ lchild,
    const u32 *__restrict__ rchild, const u8 *__restrict__ lflag, const u8 *__restrict__ rflag,
    const u32 *__restrict__ parent, u32 *__restrict__ counters, f64 scale, f64 *out) {
    u32 leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= L)
        return;
    f64 v = -1.7976931348623157e308;
    u32 a = reduc_index_map[leaf], b = reduc_index_map[leaf + 1];
    for (u32 sidx = a; sidx < b; sidx++)
        v = fmax(v, field[u64(index_map[sidx]) * fstride]);
    u32 node  = I + leaf;
    out[node] = v * scale; // max commutes with the positive scale (Solver.cpp:1345-1352)
    if (I == 0)
        return;
    node = parent[node];
    while (true) {
        __threadfence();
        u32 old = atomicAdd(&counters[node], 1u);
        if (old == 0)
            return;
        u32 l = lchild[node] + I * u32(lflag[node]);
        u32 r = rchild[node] + I * u32(rflag[node]);
        out[node] = fmax(__ldcg(&out[l]), __ldcg(&out[r]));
        if (node == 0)
            return;
        node = parent[node];
    }
}

// =============================================================================================
// host drivers
// =============================================================================================
/// Morton codes over the box [bmin, bmax] + key/value sort: t.index_map[0..M) is the Morton order of the
/// objects (shamtree/src/RadixTreeMortonBuilder.cpp:68-107, the part modules::ParticleReordering needs)
/// sort_mode RADIX: the onesweep sort on the 30 bits of a Morton code (SHAMB200_ONESWEEP=0: the three-kernel
/// LSD sort of round 1, for A / B runs)
static void sort_radix(cudaStream_t s, TreeBuffers &t, u32 M) {
    const char *e = getenv("SHAMB200_ONESWEEP");
    if (e && atoi(e) == 0)
        radix_sort_by_key(s, t.morton.p, t.index_map.p, t.morton_alt.p, t.index_alt.p, M, 32, t.radix_hist);
    else // index_map holds 0..M-1 (morton_kernel): the first pass generates it instead of reading it
        onesweep_sort_by_key(s, t.morton.p, t.index_map.p, t.morton_alt.p, t.index_alt.p, M, 30, t.radix_hist, true);
}

void morton_sort_permutation(
    cudaStream_t s, TreeBuffers &t, const f64 *d_xyz, size_t stride, u32 M, const f64 *bmin, const f64 *bmax,
    int sort_mode) {
    if (M == 0)
        return;
    t.M  = M;
    t.P2 = roundup_pow2(M);
    t.bbox.ensure(8);
    t.scalars.ensure(16);
    t.h_scalars.ensure(16);
    f64 hb[6] = {bmin[0], bmin[1], bmin[2], bmax[0], bmax[1], bmax[2]};
    f64 *hp   = reinterpret_cast<f64 *>(t.h_scalars.p + 8);
    std::memcpy(hp, hb, sizeof(hb));
    h2d_small(s, t.bbox.p, hp, sizeof(hb));
    t.morton.ensure(t.P2);
    t.index_map.ensure(t.P2);
    morton_kernel<<<grid_for(t.P2, 256), 256, 0, s>>>(d_xyz, stride, M, t.P2, t.bbox.p, t.morton.p, t.index_map.p);
    SB_COUNT_LAUNCH();
    if (sort_mode == SORT_RADIX) {
        t.morton_alt.ensure(t.P2);
        t.index_alt.ensure(t.P2);
        sort_radix(s, t, M);
    } else {
        bitonic_sort_by_key(s, t.morton.p, t.index_map.p, t.P2);
    }
    SB_LAUNCH_CHECK();
    // the box went up from the pinned staging words of `t`: they must not be rewritten (next patch) before
    // the copy has run
    SB_CUDA_CHECK(cudaStreamSynchronize(s));
}

void tree_build(
    cudaStream_t s, TreeBuffers &t, const f64 *d_xyz, size_t stride, u32 M, const f64 *bmin,
    const f64 *bmax, bool auto_bbox, u32 reduction_level, int sort_mode, f64 field_scale, DevBuf<f64> *field_out) {
    tree_build_begin(s, t, d_xyz, stride, M, bmin, bmax, auto_bbox, reduction_level, sort_mode);
    SB_CUDA_CHECK(cudaStreamSynchronize(s));
    tree_build_finish(s, t, d_xyz, stride, field_scale, field_out);
}

void tree_build_begin(
    cudaStream_t s, TreeBuffers &t, const f64 *d_xyz, size_t stride, u32 M, const f64 *bmin,
    const f64 *bmax, bool auto_bbox, u32 reduction_level, int sort_mode) {
    if (M == 0)
        throw std::invalid_argument("obj_cnt is 0, cannot build a CompressedLeafBVH");
    t.M  = M;
    t.P2 = roundup_pow2(M);
    t.bbox.ensure(8);
    t.scalars.ensure(16);
    t.h_scalars.ensure(16);
    if (auto_bbox) {
        u64 *acc = t.scalars.p + 8;
        bbox_init_kernel<<<1, 32, 0, s>>>(acc);
        SB_COUNT_LAUNCH();
        unsigned nb = (unsigned) std::min<u64>(u64(kNumSM) * 8, (u64(M) + 255) / 256);
        bbox_reduce_kernel<<<nb, 256, 0, s>>>(d_xyz, stride, M, acc);
        SB_COUNT_LAUNCH();
        bbox_finalize_kernel<<<1, 32, 0, s>>>(acc, t.bbox.p, true);
        SB_COUNT_LAUNCH();
    } else {
        f64 hb[6] = {bmin[0], bmin[1], bmin[2], bmax[0], bmax[1], bmax[2]};
        f64 *hp   = reinterpret_cast<f64 *>(t.h_scalars.p + 8);
        std::memcpy(hp, hb, sizeof(hb));
        h2d_small(s, t.bbox.p, hp, sizeof(hb));
    }
    t.morton.ensure(t.P2);
    t.index_map.ensure(t.P2);
    morton_kernel<<<grid_for(t.P2, 256), 256, 0, s>>>(d_xyz, stride, M, t.P2, t.bbox.p, t.morton.p, t.index_map.p);
    SB_COUNT_LAUNCH();
    if (sort_mode == SORT_RADIX) {
        t.morton_alt.ensure(t.P2);
        t.index_alt.ensure(t.P2);
        // only the M real keys: the padding (0xFFFFFFFF, above every 30-bit code) is already in place behind them
        sort_radix(s, t, M);
    } else {
        bitonic_sort_by_key(s, t.morton.p, t.index_map.p, t.P2);
    }
    // leaf compression
    t.split1.ensure(M);
    t.split2.ensure(M);
    // (a version that keeps the split table and the reduction iterations of a tile in shared memory was measured
    // at 0.97 ms against 0.41 ms for these four streaming launches at 17 M keys: profiles/README.md)
    split_table_kernel<<<grid_for(M, 256), 256, 0, s>>>(t.morton.p, M, t.split1.p);
    SB_COUNT_LAUNCH();
    u8 *cur = t.split1.p, *oth = t.split2.p;
    for (u32 it = 1; it <= reduction_level; it++) {
        reduction_iteration_kernel<<<grid_for(M, 256), 256, 0, s>>>(t.morton.p, M, cur, oth);
        SB_COUNT_LAUNCH();
        std::swap(cur, oth);
    }
    t.scan_out.ensure(M);
    exclusive_scan<u8>(s, cur, t.scan_out.p, M, t.scan_tmp, t.scalars.p);
    t.reduc_index_map.ensure(size_t(M) + 2);
    t.reduced_morton.ensure(M);
    compact_leaves_kernel<<<grid_for(M, 256), 256, 0, s>>>(
        t.morton.p, M, cur, t.scan_out.p, t.scalars.p, t.reduc_index_map.p, t.reduced_morton.p);
    SB_COUNT_LAUNCH();
    // the leaf count (and the bbox) are needed on the host
    d2h_small(s, t.h_scalars.p, t.scalars.p, sizeof(u64));
    d2h_small(s, t.h_scalars.p + 1, t.bbox.p, 6 * sizeof(f64));
    SB_LAUNCH_CHECK();
}

void tree_build_finish(
    cudaStream_t s, TreeBuffers &t, const f64 *d_xyz, size_t stride, f64 field_scale, DevBuf<f64> *field_out) {
    t.L = u32(t.h_scalars.p[0]);
    std::memcpy(t.bmin, t.h_scalars.p + 1, 3 * sizeof(f64));
    std::memcpy(t.bmax, t.h_scalars.p + 4, 3 * sizeof(f64));
    if (t.L == 0)
        throw std::runtime_error("0 leaf tree cannot exists");
    t.I = t.L - 1;
    size_t ni = t.I ? t.I : 1;
    t.lchild.ensure(ni);
    t.rchild.ensure(ni);
    t.endrange.ensure(ni);
    t.lflag.ensure(ni);
    t.rflag.ensure(ni);
    t.parent.ensure(size_t(t.I) + t.L);
    t.counters.ensure(ni);
    if (t.I) {
        karras_kernel<<<grid_for(t.I, 256), 256, 0, s>>>(
            t.reduced_morton.p, t.I, t.lchild.p, t.rchild.p, t.lflag.p, t.rflag.p, t.endrange.p, t.parent.p);
        SB_COUNT_LAUNCH();
        SB_CUDA_CHECK(cudaMemsetAsync(t.counters.p, 0, size_t(t.I) * sizeof(u32), s));
    }
    size_t tot = size_t(t.I) + t.L;
    t.aabb_min.ensure(tot * 3);
    t.aabb_max.ensure(tot * 3);
    if (field_out) {
        if (stride < 4)
            throw std::invalid_argument("tree_build: the fused field maximum needs (x, y, z, f) records");
        field_out->ensure(tot);
        leaf_aabb_propagate_kernel<true><<<grid_for(t.L, 128), 128, 0, s>>>(
            d_xyz, stride, t.index_map.p, t.reduc_index_map.p, t.L, t.I, t.lchild.p, t.rchild.p, t.lflag.p,
            t.rflag.p, t.parent.p, t.counters.p, t.aabb_min.p, t.aabb_max.p, field_scale, field_out->p);
    } else {
        leaf_aabb_propagate_kernel<false><<<grid_for(t.L, 128), 128, 0, s>>>(
            d_xyz, stride, t.index_map.p, t.reduc_index_map.p, t.L, t.I, t.lchild.p, t.rchild.p, t.lflag.p,
            t.rflag.p, t.parent.p, t.counters.p, t.aabb_min.p, t.aabb_max.p, 0., nullptr);
    }
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

void tree_field_max(cudaStream_t s, TreeBuffers &t, const f64 *d_field, f64 scale, f64 *d_out, size_t field_stride) {
    if (t.I)
        SB_CUDA_CHECK(cudaMemsetAsync(t.counters.p, 0, size_t(t.I) * sizeof(u32), s));
    leaf_field_max_propagate_kernel<<<grid_for(t.L, 128), 128, 0, s>>>(
        d_field, field_stride, t.index_map.p, t.reduc_index_map.p, t.L, t.I, t.lchild.p, t.rchild.p, t.lflag.p,
        t.rflag.p, t.parent.p, t.counters.p, scale, d_out);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

} // namespace sb
