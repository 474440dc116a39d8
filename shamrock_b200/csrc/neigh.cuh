// neigh.cuh — neighbour cache buffers (shamrock::tree::ObjectCache) and builder.
#pragma once
#include "tree.cuh"

namespace sb {

struct NeighBuffers {
    u32 N = 0; ///< objects with a list (real particles)
    u32 K = 0; ///< total number of neighbours (u32 like the reference)
    DevBuf<u32> leaf_cnt, leaf_scanned, leaf_list, owner;
    DevBuf<u32> cnt, scanned, list;
    DevBuf<u32> scan_tmp;
    DevBuf<u64> scalars;
    PinnedBuf<u64> h_scalars;
};

void neigh_cache_build(
    cudaStream_t s, const TreeBuffers &tb, NeighBuffers &nb, const f64 *d_xyz, size_t stride_dbl,
    const f64 *d_hpart, const f64 *d_rint, u32 N, f64 Rkern, f64 h_tolerance, bool two_stage,
    size_t h_stride = 1);

} // namespace sb
