// tree.cuh — device-side compressed-leaf BVH (shamtree::CompressedLeafBVH<u32,f64_3,3>) for sm_100a.
#pragma once
#include "common.cuh"
#include "primitives.cuh"

namespace sb {

/// All device arrays of one tree (grow-only, recycled between rebuilds like the reference's
/// move-in "cache" arguments, CompressedLeafBVH.cpp:45-66).
struct TreeBuffers {
    u32 M = 0, P2 = 0, L = 0, I = 0;
    f64 bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};
    DevBuf<u32> morton, index_map;     // [P2]
    DevBuf<u32> morton_alt, index_alt; // [P2] radix ping-pong
    DevBuf<u8> split1, split2;         // [M]
    DevBuf<u32> scan_out, scan_tmp;    // [M], blocks
    DevBuf<u32> reduc_index_map;       // [L+2]
    DevBuf<u32> reduced_morton;        // [L]
    DevBuf<u32> lchild, rchild, endrange, parent; // [I], parent [I+L]
    DevBuf<u8> lflag, rflag;           // [I]
    DevBuf<f64> aabb_min, aabb_max;    // [(I+L)*3]
    DevBuf<u32> counters;              // [I]
    DevBuf<f64> bbox;                  // 6 doubles + transform (bmin[3], bmax[3])
    DevBuf<u64> scalars;               // device scalars (totals)
    DevBuf<u32> radix_hist;            // radix histograms
    PinnedBuf<u64> h_scalars;
};

enum SortMode { SORT_BITONIC = 0, SORT_RADIX = 1 };

/// d_bbox (device, 6 doubles: bmin, bmax) may be null → bmin/bmax host values are used.
void tree_build(
    cudaStream_t s, TreeBuffers &t, const f64 *d_xyz, size_t stride_dbl, u32 M, const f64 *bmin,
    const f64 *bmax, bool auto_bbox, u32 reduction_level, int sort_mode, f64 field_scale = 0.,
    DevBuf<f64> *field_out = nullptr);
/// the same in two pieces around ONE stream synchronisation of the caller (several trees: enqueue all, synchronise
/// once, finish all): `begin` runs up to the leaf compression and starts the read-back of the leaf count,
/// `finish` (after the synchronisation) builds the Karras tree and the boxes
void tree_build_begin(
    cudaStream_t s, TreeBuffers &t, const f64 *d_xyz, size_t stride_dbl, u32 M, const f64 *bmin,
    const f64 *bmax, bool auto_bbox, u32 reduction_level, int sort_mode);
void tree_build_finish(
    cudaStream_t s, TreeBuffers &t, const f64 *d_xyz, size_t stride_dbl, f64 field_scale = 0.,
    DevBuf<f64> *field_out = nullptr);
// field_out (optional; stride_dbl >= 4): the objects are (x, y, z, f) records and the AABB pass also leaves
// field_out[node] = field_scale * max f of the node's objects ([I+L]) — tree_field_max without its own pass

/// Morton codes over [bmin, bmax] + sort only: t.index_map[0..M) = Morton order of the objects
void morton_sort_permutation(
    cudaStream_t s, TreeBuffers &t, const f64 *d_xyz, size_t stride_dbl, u32 M, const f64 *bmin, const f64 *bmax,
    int sort_mode);

/// out[node] = scale * max over the node's objects of field[obj]   ([I+L])
void tree_field_max(cudaStream_t s, TreeBuffers &t, const f64 *d_field, f64 scale, f64 *d_out, size_t field_stride = 1);

/// sort (key,value) pairs of length P2 (power of two) exactly like the reference's bitonic network
void bitonic_sort_by_key(cudaStream_t s, u32 *keys, u32 *vals, u32 len);
/// stable LSD radix sort on `bits` key bits (CUB-free); result in keys/vals (uses alt buffers)
void radix_sort_by_key(
    cudaStream_t s, u32 *keys, u32 *vals, u32 *keys_alt, u32 *vals_alt, u32 len, int bits,
    DevBuf<u32> &hist);

/// the same sort in one histogram kernel + one decoupled-look-back kernel per digit (no host synchronisation)
void onesweep_sort_by_key(
    cudaStream_t s, u32 *keys, u32 *vals, u32 *keys_alt, u32 *vals_alt, u32 len, int bits, DevBuf<u32> &work,
    bool iota_values);

inline u32 roundup_pow2(u32 v) {
    if (v <= 1)
        return 1;
    if ((v & (v - 1)) == 0)
        return v;
    return 1u << (32 - __builtin_clz(v));
}

} // namespace sb
