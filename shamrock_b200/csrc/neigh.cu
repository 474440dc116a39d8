// neigh.cu — per-patch neighbour cache (shamrock::tree::ObjectCache) on sm_100a.
//
// Restates shammodels/sph/src/modules/NeighbourCache.cpp:223-604 (two-stage: leaf→leaf list, owner
// leaf by point location, per-particle scan of the neighbouring leaves) and :30-220 (one-stage tree
// walk with sph_radix_cell_crit, shamtree/include/shamtree/RadixTree.hpp:802-816); traversal order
// of shamtree/include/shamtree/KarrasTreeTraverser.hpp:71-118 (stack DFS, left child first).
// Compiled with -fmad=false: the accept test `r² > (h·tol)²·R²` must round exactly like the
// reference's separate multiplications (bit-exact neighbour lists).
//
// Work distribution (B200): one thread per object *in sorted-Morton order* (thread r handles
// object sort_index_map[r]) so that the 32 lanes of a warp share the same few leaves: the
// candidate tiles they gather are the same cache lines.
#include "neigh.cuh"

namespace sb {

__device__ __forceinline__ bool cella_neigh_b(
    f64 ax0, f64 ay0, f64 az0, f64 ax1, f64 ay1, f64 az1, f64 bx0, f64 by0, f64 bz0, f64 bx1,
    f64 by1, f64 bz1) {
    return (fmax(ax0, bx0) <= fmin(ax1, bx1)) && (fmax(ay0, by0) <= fmin(ay1, by1))
           && (fmax(az0, bz0) <= fmin(az1, bz1));
}

struct TreeView {
    const u32 *lchild, *rchild;
    const u8 *lflag, *rflag;
    const f64 *aabb_min, *aabb_max;
    const u32 *index_map, *reduc_index_map;
    u32 I, L;
};

constexpr int STACK_DEPTH = 31; // MortonCodes<u32,3>::significant_bits + 1

/// generic stack traversal; cond(node) -> bool, on_leaf(node)
template<class Cond, class OnLeaf>
__device__ __forceinline__ void rtree_for(const TreeView &t, Cond cond, OnLeaf on_leaf) {
    u32 stack[STACK_DEPTH];
    u32 cursor    = STACK_DEPTH - 1;
    stack[cursor] = 0;
    while (cursor < STACK_DEPTH) {
        u32 cur = stack[cursor];
        cursor++;
        if (cond(cur)) {
            if (cur >= t.I) {
                on_leaf(cur);
            } else {
                u32 l = t.lchild[cur] + t.I * u32(t.lflag[cur]);
                u32 r = t.rchild[cur] + t.I * u32(t.rflag[cur]);
                stack[--cursor] = r;
                stack[--cursor] = l;
            }
        }
    }
}

// ---- stage 1: leaf -> leaf ---------------------------------------------------------------------
template<bool FILL>
__global__ void __launch_bounds__(128) leaf_leaf_kernel(
    TreeView t, const f64 *__restrict__ rint, f64 Rkern, u32 *__restrict__ cnt_out,
    const u32 *__restrict__ scanned, u32 *__restrict__ list) {
    u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= t.L)
        return;
    u32 na     = t.I + g;
    f64 a_rint = rint[na] * Rkern;
    f64 a0x = t.aabb_min[3 * u64(na)], a0y = t.aabb_min[3 * u64(na) + 1], a0z = t.aabb_min[3 * u64(na) + 2];
    f64 a1x = t.aabb_max[3 * u64(na)], a1y = t.aabb_max[3 * u64(na) + 1], a1z = t.aabb_max[3 * u64(na) + 2];
    f64 e0x = a0x - a_rint, e0y = a0y - a_rint, e0z = a0z - a_rint;
    f64 e1x = a1x + a_rint, e1y = a1y + a_rint, e1z = a1z + a_rint;
    u32 cnt = FILL ? scanned[g] : 0u;
    rtree_for(
        t,
        [&](u32 node) {
            f64 r   = rint[node] * Rkern;
            f64 n0x = t.aabb_min[3 * u64(node)], n0y = t.aabb_min[3 * u64(node) + 1], n0z = t.aabb_min[3 * u64(node) + 2];
            f64 n1x = t.aabb_max[3 * u64(node)], n1y = t.aabb_max[3 * u64(node) + 1], n1z = t.aabb_max[3 * u64(node) + 2];
            return cella_neigh_b(a0x, a0y, a0z, a1x, a1y, a1z, n0x - r, n0y - r, n0z - r, n1x + r, n1y + r, n1z + r)
                   || cella_neigh_b(e0x, e0y, e0z, e1x, e1y, e1z, n0x, n0y, n0z, n1x, n1y, n1z);
        },
        [&](u32 leaf_b) {
            if (FILL)
                list[cnt] = leaf_b;
            cnt++;
        });
    if (!FILL)
        cnt_out[g] = cnt;
}

// ---- stage 2a: owner leaf of each object ---------------------------------------------------------
__global__ void __launch_bounds__(128) leaf_owner_kernel(
    TreeView t, const f64 *__restrict__ xyz, size_t stride, u32 M, u32 N, u32 *__restrict__ owner) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M)
        return;
    u32 id_a = t.index_map[r];
    if (id_a >= N)
        return;
    f64 x = xyz[u64(id_a) * stride], y = xyz[u64(id_a) * stride + 1], z = xyz[u64(id_a) * stride + 2];
    u32 found = 0x7fffffffu;
    rtree_for(
        t,
        [&](u32 node) {
            return (t.aabb_min[3 * u64(node)] <= x) && (x <= t.aabb_max[3 * u64(node)])
                   && (t.aabb_min[3 * u64(node) + 1] <= y) && (y <= t.aabb_max[3 * u64(node) + 1])
                   && (t.aabb_min[3 * u64(node) + 2] <= z) && (z <= t.aabb_max[3 * u64(node) + 2]);
        },
        [&](u32 leaf_b) { found = leaf_b - t.I; });
    owner[id_a] = found;
}

// ---- stage 2b: objects of the neighbouring leaves -------------------------------------------------
template<bool FILL>
__global__ void __launch_bounds__(128) particle_neigh_2stage_kernel(
    TreeView t, const f64 *__restrict__ xyz, size_t stride, const f64 *__restrict__ hpart, size_t hstride, u32 M,
    u32 N, const u32 *__restrict__ owner, const u32 *__restrict__ leaf_cnt,
    const u32 *__restrict__ leaf_scanned, const u32 *__restrict__ leaf_list, f64 Rker2,
    f64 h_tolerance, u32 *__restrict__ cnt_out, const u32 *__restrict__ scanned, u32 *__restrict__ list) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M)
        return;
    u32 id_a = t.index_map[r];
    if (id_a >= N)
        return;
    f64 rint_a = hpart[u64(id_a) * hstride] * h_tolerance;
    f64 lim_a  = rint_a * rint_a * Rker2;
    f64 ax = xyz[u64(id_a) * stride], ay = xyz[u64(id_a) * stride + 1], az = xyz[u64(id_a) * stride + 2];
    u32 cnt = FILL ? scanned[id_a] : 0u;
    u32 own = owner[id_a];
    u32 s0 = leaf_scanned[own], s1 = s0 + leaf_cnt[own];
    for (u32 k = s0; k < s1; k++) {
        u32 leaf_b = leaf_list[k] - t.I;
        u32 p0 = t.reduc_index_map[leaf_b], p1 = t.reduc_index_map[leaf_b + 1];
        for (u32 sidx = p0; sidx < p1; sidx++) {
            u32 id_b = t.index_map[sidx];
            f64 dx = ax - xyz[u64(id_b) * stride], dy = ay - xyz[u64(id_b) * stride + 1],
                dz = az - xyz[u64(id_b) * stride + 2];
            f64 rab2   = dx * dx + dy * dy + dz * dz;
            f64 rint_b = hpart[u64(id_b) * hstride] * h_tolerance;
            bool no_interact = rab2 > lim_a && rab2 > rint_b * rint_b * Rker2;
            if (!no_interact) {
                if (FILL)
                    list[cnt] = id_b;
                cnt++;
            }
        }
    }
    if (!FILL)
        cnt_out[id_a] = cnt;
}

// ---- one-stage variant ---------------------------------------------------------------------------
template<bool FILL>
__global__ void __launch_bounds__(128) particle_neigh_1stage_kernel(
    TreeView t, const f64 *__restrict__ xyz, size_t stride, const f64 *__restrict__ hpart, size_t hstride,
    const f64 *__restrict__ rint, u32 M, u32 N, f64 Rkern, f64 h_tolerance,
    u32 *__restrict__ cnt_out, const u32 *__restrict__ scanned, u32 *__restrict__ list) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M)
        return;
    u32 id_a = t.index_map[r];
    if (id_a >= N)
        return;
    const f64 Rker2 = Rkern * Rkern;
    f64 rint_a = hpart[u64(id_a) * hstride] * h_tolerance;
    f64 lim_a  = rint_a * rint_a * Rker2;
    f64 ax = xyz[u64(id_a) * stride], ay = xyz[u64(id_a) * stride + 1], az = xyz[u64(id_a) * stride + 2];
    f64 ra  = rint_a * Rkern;
    f64 b0x = ax - ra, b0y = ay - ra, b0z = az - ra, b1x = ax + ra, b1y = ay + ra, b1z = az + ra;
    u32 cnt = FILL ? scanned[id_a] : 0u;
    rtree_for(
        t,
        [&](u32 node) {
            f64 rr  = rint[node] * Rkern;
            f64 n0x = t.aabb_min[3 * u64(node)], n0y = t.aabb_min[3 * u64(node) + 1], n0z = t.aabb_min[3 * u64(node) + 2];
            f64 n1x = t.aabb_max[3 * u64(node)], n1y = t.aabb_max[3 * u64(node) + 1], n1z = t.aabb_max[3 * u64(node) + 2];
            return cella_neigh_b(b0x, b0y, b0z, b1x, b1y, b1z, n0x, n0y, n0z, n1x, n1y, n1z)
                   || cella_neigh_b(ax, ay, az, ax, ay, az, n0x - rr, n0y - rr, n0z - rr, n1x + rr, n1y + rr, n1z + rr);
        },
        [&](u32 leaf) {
            u32 lb = leaf - t.I;
            u32 p0 = t.reduc_index_map[lb], p1 = t.reduc_index_map[lb + 1];
            for (u32 sidx = p0; sidx < p1; sidx++) {
                u32 id_b = t.index_map[sidx];
                f64 dx = ax - xyz[u64(id_b) * stride], dy = ay - xyz[u64(id_b) * stride + 1],
                    dz = az - xyz[u64(id_b) * stride + 2];
                f64 rab2   = dx * dx + dy * dy + dz * dz;
                f64 rint_b = hpart[u64(id_b) * hstride] * h_tolerance;
                bool no_interact = rab2 > lim_a && rab2 > rint_b * rint_b * Rker2;
                if (!no_interact) {
                    if (FILL)
                        list[cnt] = id_b;
                    cnt++;
                }
            }
        });
    if (!FILL)
        cnt_out[id_a] = cnt;
}

void neigh_cache_build(
    cudaStream_t s, const TreeBuffers &tb, NeighBuffers &nb, const f64 *d_xyz, size_t stride,
    const f64 *d_hpart, const f64 *d_rint, u32 N, f64 Rkern, f64 h_tolerance, bool two_stage,
    size_t h_stride) {
    TreeView t{tb.lchild.p,   tb.rchild.p,    tb.lflag.p,         tb.rflag.p, tb.aabb_min.p,
               tb.aabb_max.p, tb.index_map.p, tb.reduc_index_map.p, tb.I,       tb.L};
    const u32 M = tb.M;
    nb.N        = N;
    nb.scalars.ensure(4);
    nb.h_scalars.ensure(4);
    nb.cnt.ensure(N);
    nb.scanned.ensure(N);
    const f64 Rker2 = Rkern * Rkern;
    if (two_stage) {
        nb.leaf_cnt.ensure(tb.L);
        nb.leaf_scanned.ensure(tb.L);
        leaf_leaf_kernel<false><<<grid_for(tb.L, 128), 128, 0, s>>>(t, d_rint, Rkern, nb.leaf_cnt.p, nullptr, nullptr);
        SB_COUNT_LAUNCH();
        exclusive_scan<u32>(s, nb.leaf_cnt.p, nb.leaf_scanned.p, tb.L, nb.scan_tmp, nb.scalars.p);
        // owner search does not depend on the leaf list: enqueue it before the size read-back
        nb.owner.ensure(N);
        leaf_owner_kernel<<<grid_for(M, 128), 128, 0, s>>>(t, d_xyz, stride, M, N, nb.owner.p);
        SB_COUNT_LAUNCH();
        SB_CUDA_CHECK(cudaMemcpyAsync(nb.h_scalars.p, nb.scalars.p, sizeof(u64), cudaMemcpyDeviceToHost, s));
        SB_CUDA_CHECK(cudaStreamSynchronize(s));
        u64 leaf_total = nb.h_scalars.p[0];
        if (leaf_total > 0xFFFFFFFFull)
            throw std::overflow_error("leaf neighbour count overflows u32");
        nb.leaf_list.ensure(leaf_total, 1.1);
        leaf_leaf_kernel<true><<<grid_for(tb.L, 128), 128, 0, s>>>(t, d_rint, Rkern, nullptr, nb.leaf_scanned.p, nb.leaf_list.p);
        SB_COUNT_LAUNCH();
        particle_neigh_2stage_kernel<false><<<grid_for(M, 128), 128, 0, s>>>(
            t, d_xyz, stride, d_hpart, h_stride, M, N, nb.owner.p, nb.leaf_cnt.p, nb.leaf_scanned.p, nb.leaf_list.p,
            Rker2, h_tolerance, nb.cnt.p, nullptr, nullptr);
        SB_COUNT_LAUNCH();
    } else {
        particle_neigh_1stage_kernel<false><<<grid_for(M, 128), 128, 0, s>>>(
            t, d_xyz, stride, d_hpart, h_stride, d_rint, M, N, Rkern, h_tolerance, nb.cnt.p, nullptr, nullptr);
        SB_COUNT_LAUNCH();
    }
    exclusive_scan<u32>(s, nb.cnt.p, nb.scanned.p, N, nb.scan_tmp, nb.scalars.p + 1);
    SB_CUDA_CHECK(cudaMemcpyAsync(nb.h_scalars.p + 1, nb.scalars.p + 1, sizeof(u64), cudaMemcpyDeviceToHost, s));
    SB_CUDA_CHECK(cudaStreamSynchronize(s));
    u64 total = nb.h_scalars.p[1];
    if (total > 0xFFFFFFFFull)
        throw std::overflow_error(
            "neighbour count overflows u32 (sum_neigh_cnt is u32 in the reference, TreeTraversal.hpp:378): "
            "use more / smaller patches");
    nb.K = u32(total);
    nb.list.ensure(total, 1.05);
    if (two_stage) {
        particle_neigh_2stage_kernel<true><<<grid_for(M, 128), 128, 0, s>>>(
            t, d_xyz, stride, d_hpart, h_stride, M, N, nb.owner.p, nb.leaf_cnt.p, nb.leaf_scanned.p, nb.leaf_list.p,
            Rker2, h_tolerance, nullptr, nb.scanned.p, nb.list.p);
    } else {
        particle_neigh_1stage_kernel<true><<<grid_for(M, 128), 128, 0, s>>>(
            t, d_xyz, stride, d_hpart, h_stride, d_rint, M, N, Rkern, h_tolerance, nullptr, nb.scanned.p, nb.list.p);
    }
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

} // namespace sb
