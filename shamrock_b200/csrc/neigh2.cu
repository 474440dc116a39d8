// neigh2.cu — B200 neighbour search (see neigh2.cuh).  Compiled with -fmad=false: the accept test
// `r² > (h·tol)²·R²` rounds exactly like the reference's separate multiplications, so the lists are
// bit-identical to shammodels/sph/src/modules/NeighbourCache.cpp:223-604 (two-stage search):
//   stage 1  leaf a ↔ node b :  cella_neigh_b(a, b ⊕ rint_b·R) || cella_neigh_b(a ⊕ rint_a·R, b)
//            (NeighbourCache.cpp:286-320, shamtree/include/shamtree/kernels/geometry_utils.hpp:126-135)
//   stage 2  particle a ↔ particles of the neighbouring leaves, accept iff
//            !(r² > (h_a·tol)²·R² && r² > (h_b·tol)²·R²)          (NeighbourCache.cpp:482-520)
// and every list is ordered by ascending rank in the sorted Morton array.  A particle's
// owner leaf is the leaf holding its rank: leaves are disjoint Morton-prefix boxes, so the reference's
// point-location descent (NeighbourCache.cpp:415-447) finds exactly that leaf.
//
// B200 mapping: particles are gathered into Morton order first, so a leaf's particles and every
// candidate run are contiguous (coalesced 32-byte records).  Stage 1 walks the packed 64-byte nodes
// once per GROUP of 8 consecutive leaves, one warp per group, 32 frontier nodes per step, and emits the
// group's candidate leaves with the mask of the members each one hits (exact per-leaf test).  Stage 2 is
// warp-per-leaf with the 32 lanes holding 32 CANDIDATES; the leaf's particles are broadcast from shared
// memory, each (particle, chunk) produces one ballot word.  The fill pass replays the ballots only.
#include "neigh2.cuh"
#include <cstdlib>
#include <algorithm>

namespace sb {

// ---------------------------------------------------------------------------------------------
// sorted storage
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sort_gather_kernel(
    const Pack4 *__restrict__ A, const u32 *__restrict__ index_map, u32 M, u32 N, Pack4 *__restrict__ SA,
    u32 *__restrict__ inv_map, u8 *__restrict__ real_flag) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == M)
        real_flag[M] = 0;
    if (r >= M)
        return;
    u32 id       = index_map[r];
    const double2 *q = reinterpret_cast<const double2 *>(A + id);
    double2 lo = __ldg(q), hi = __ldg(q + 1);
    double2 *o = reinterpret_cast<double2 *>(SA + r);
    o[0]       = lo;
    o[1]       = hi;
    inv_map[id]  = r;
    real_flag[r] = id < N ? 1 : 0;
}
__global__ void __launch_bounds__(256) sort_gather_strided_kernel(
    const f64 *__restrict__ xyz, size_t stride, const f64 *__restrict__ h, size_t hstride,
    const u32 *__restrict__ index_map, u32 M, u32 N, Pack4 *__restrict__ SA, u32 *__restrict__ inv_map,
    u8 *__restrict__ real_flag) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == M)
        real_flag[M] = 0;
    if (r >= M)
        return;
    u32 id = index_map[r];
    SA[r]  = Pack4{xyz[u64(id) * stride], xyz[u64(id) * stride + 1], xyz[u64(id) * stride + 2], h[u64(id) * hstride]};
    inv_map[id]  = r;
    real_flag[r] = id < N ? 1 : 0;
}
__global__ void __launch_bounds__(256) slot_rank_kernel(
    const u8 *__restrict__ real_flag, const u32 *__restrict__ real_prefix, u32 M, u32 *__restrict__ slot_rank) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < M && real_flag[r])
        slot_rank[real_prefix[r]] = r;
}

static void finish_sorted(cudaStream_t s, SearchBuffers &sb) {
    sb.scalars.ensure(8);
    sb.h_scalars.ensure(8);
    sb.real_prefix.ensure(size_t(sb.M) + 1);
    exclusive_scan<u8>(s, sb.real_flag.p, sb.real_prefix.p, u64(sb.M) + 1, sb.scan_tmp, sb.scalars.p);
    sb.slot_rank.ensure(sb.N);
    slot_rank_kernel<<<grid_for(sb.M, 256), 256, 0, s>>>(sb.real_flag.p, sb.real_prefix.p, sb.M, sb.slot_rank.p);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

void search_prepare_sorted(cudaStream_t s, const TreeBuffers &tb, SearchBuffers &sb, const Pack4 *A, u32 N) {
    sb.M = tb.M;
    sb.N = N;
    sb.L = tb.L;
    sb.I = tb.I;
    sb.exported = false;
    sb.SA.ensure(sb.M, 1.1);
    sb.inv_map.ensure(sb.M, 1.1);
    sb.real_flag.ensure(size_t(sb.M) + 1, 1.1);
    sort_gather_kernel<<<grid_for(u64(sb.M) + 1, 256), 256, 0, s>>>(
        A, tb.index_map.p, sb.M, N, sb.SA.p, sb.inv_map.p, sb.real_flag.p);
    SB_COUNT_LAUNCH();
    finish_sorted(s, sb);
}
void search_prepare_sorted_strided(
    cudaStream_t s, const TreeBuffers &tb, SearchBuffers &sb, const f64 *xyz, size_t stride, const f64 *h,
    size_t hstride, u32 N) {
    sb.M = tb.M;
    sb.N = N;
    sb.L = tb.L;
    sb.I = tb.I;
    sb.exported = false;
    sb.SA.ensure(sb.M, 1.1);
    sb.inv_map.ensure(sb.M, 1.1);
    sb.real_flag.ensure(size_t(sb.M) + 1, 1.1);
    sort_gather_strided_kernel<<<grid_for(u64(sb.M) + 1, 256), 256, 0, s>>>(
        xyz, stride, h, hstride, tb.index_map.p, sb.M, N, sb.SA.p, sb.inv_map.p, sb.real_flag.p);
    SB_COUNT_LAUNCH();
    finish_sorted(s, sb);
}

// ---------------------------------------------------------------------------------------------
// packed nodes
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_nodes_kernel(
    u32 I, u32 L, const f64 *__restrict__ aabb_min, const f64 *__restrict__ aabb_max, const f64 *__restrict__ rint,
    const u32 *__restrict__ lchild, const u32 *__restrict__ rchild, const u8 *__restrict__ lflag,
    const u8 *__restrict__ rflag, const u32 *__restrict__ reduc_index_map, NodePack *__restrict__ nodes) {
    u32 n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= I + L)
        return;
    NodePack p;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        p.lo[c] = aabb_min[3 * u64(n) + c];
        p.hi[c] = aabb_max[3 * u64(n) + c];
    }
    p.rint = rint[n];
    if (n < I) {
        p.left  = lchild[n] + I * u32(lflag[n]);
        p.right = rchild[n] + I * u32(rflag[n]);
    } else {
        p.left  = reduc_index_map[n - I];
        p.right = reduc_index_map[n - I + 1];
    }
    nodes[n] = p;
}

// ---------------------------------------------------------------------------------------------
// stage 1: candidate leaves of every GROUP of GL consecutive leaves (one warp-cooperative walk per group)
// ---------------------------------------------------------------------------------------------
struct NodeRegs {
    f64 lo0, lo1, lo2, hi0, hi1, hi2, rint;
    u32 left, right;
};
__device__ __forceinline__ NodeRegs load_node(const NodePack *p) {
    const Pack4 *q = reinterpret_cast<const Pack4 *>(p); // two 256-bit loads
    Pack4 a = ld4(q), b = ld4(q + 1);
    NodeRegs n;
    n.lo0 = a.a, n.lo1 = a.b, n.lo2 = a.c, n.hi0 = a.d, n.hi1 = b.a, n.hi2 = b.b, n.rint = b.c;
    u64 ch  = (u64) __double_as_longlong(b.d);
    n.left  = u32(ch & 0xffffffffull);
    n.right = u32(ch >> 32);
    return n;
}

constexpr int WALK_WARPS = 1; // one walk per block: walks differ in length, a block would wait for its slowest warp

/// box of one leaf of the group and the same box grown by its own interaction radius
struct LeafBox {
    f64 lo[3], hi[3], e0[3], e1[3];
};

__device__ __forceinline__ f64 warp_min8(f64 v) { // lanes >= 8 hold the neutral element
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1)
        v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return __shfl_sync(0xffffffffu, v, 0);
}
__device__ __forceinline__ f64 warp_max8(f64 v) {
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1)
        v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return __shfl_sync(0xffffffffu, v, 0);
}

constexpr u32 TOP_CAP_MAX = 512; ///< start-frontier entries per super-group: storage
/// groups per super-group sharing the top of their walks, and the frontier length the top walk hands over
/// (it stops once half of it is reached).  SHAMB200_SUPER / SHAMB200_TOPCAP override the defaults (tuning).
struct TopCfg {
    u32 super, cap;
};
static TopCfg top_cfg() {
    TopCfg t{64u, 128u}; // read at every search: scripts/tune.py switches them inside one process
    if (const char *e = getenv("SHAMB200_SUPER"))
        t.super = std::max(1, atoi(e));
    if (const char *e = getenv("SHAMB200_TOPCAP"))
        t.cap = std::min<u32>(TOP_CAP_MAX, std::max(4, atoi(e)));
    return t;
}

__device__ __forceinline__ f64 warp_min32(f64 v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ f64 warp_max32(f64 v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

/// The first ~15 levels of the walks of neighbouring groups visit the same nodes, one dependent memory
/// latency per level.  One warp per super-group of 64 groups descends from the root with the union of all
/// its leaves (a superset of every member group's test, so nothing a group needs is pruned) until the
/// frontier holds 64 nodes, and leaves that frontier (node ids, left to right, untested) as the starting
/// point of the 64 group walks.
__global__ void __launch_bounds__(32) top_walk_kernel(
    const NodePack *__restrict__ nodes, u32 I, u32 L, const u32 *__restrict__ real_prefix, f64 Rkern, u32 SUPER,
    u32 TOP_CAP, u32 *__restrict__ top_front, u32 *__restrict__ top_count) {
    __shared__ u32 fa[TOP_CAP_MAX], fb[TOP_CAP_MAX];
    const int lane = threadIdx.x;
    const u32 lt   = (1u << lane) - 1u;
    const u32 sg   = blockIdx.x;
    const u32 leaf0 = sg * SUPER * GL, leaf1 = min(L, leaf0 + SUPER * GL);
    const f64 inf = __longlong_as_double(0x7ff0000000000000ll);
    f64 l0 = inf, l1 = inf, l2 = inf, h0 = -inf, h1 = -inf, h2 = -inf;
    f64 p0 = inf, p1 = inf, p2 = inf, q0 = -inf, q1 = -inf, q2 = -inf;
    for (u32 leaf = leaf0 + lane; leaf < leaf1; leaf += 32) {
        NodeRegs a = load_node(nodes + I + leaf);
        if (real_prefix[a.right] == real_prefix[a.left])
            continue;
        f64 ar = a.rint * Rkern;
        l0 = fmin(l0, a.lo0), l1 = fmin(l1, a.lo1), l2 = fmin(l2, a.lo2);
        h0 = fmax(h0, a.hi0), h1 = fmax(h1, a.hi1), h2 = fmax(h2, a.hi2);
        p0 = fmin(p0, a.lo0 - ar), p1 = fmin(p1, a.lo1 - ar), p2 = fmin(p2, a.lo2 - ar);
        q0 = fmax(q0, a.hi0 + ar), q1 = fmax(q1, a.hi1 + ar), q2 = fmax(q2, a.hi2 + ar);
    }
    l0 = warp_min32(l0), l1 = warp_min32(l1), l2 = warp_min32(l2);
    h0 = warp_max32(h0), h1 = warp_max32(h1), h2 = warp_max32(h2);
    p0 = warp_min32(p0), p1 = warp_min32(p1), p2 = warp_min32(p2);
    q0 = warp_max32(q0), q1 = warp_max32(q1), q2 = warp_max32(q2);
    if (!(l0 <= h0)) { // no real particle below this super-group: its groups return at once
        if (lane == 0)
            top_count[sg] = 0;
        return;
    }
    u32 *cur = fa, *nxt = fb;
    if (lane == 0)
        cur[0] = 0u;
    u32 ncur = 1;
    for (;;) {
        __syncwarp();
        u32 nn       = 0;
        bool pending = false;
        for (u32 base = 0; base < ncur; base += 32) {
            const u32 k = base + lane;
            u32 emit = 0, o0 = 0, o1 = 0;
            if (k < ncur) {
                const u32 id = cur[k];
                if (id >= I) { // leaves are left to the group walks
                    emit = 1;
                    o0   = id;
                } else {
                    NodeRegs n = load_node(nodes + id);
                    f64 r      = n.rint * Rkern;
                    f64 x1 = n.hi0 + r, y1 = n.hi1 + r, z1 = n.hi2 + r;
                    f64 x0 = n.lo0 - r, y0 = n.lo1 - r, z0 = n.lo2 - r;
                    bool hit = ((l0 <= x1) & (x0 <= h0) & (l1 <= y1) & (y0 <= h1) & (l2 <= z1) & (z0 <= h2))
                               | ((p0 <= n.hi0) & (n.lo0 <= q0) & (p1 <= n.hi1) & (n.lo1 <= q1) & (p2 <= n.hi2)
                                  & (n.lo2 <= q2));
                    if (hit) {
                        emit    = 2;
                        o0      = n.left;
                        o1      = n.right;
                        pending = true;
                    }
                }
            }
            const u32 b1 = __ballot_sync(0xffffffffu, emit >= 1), b2 = __ballot_sync(0xffffffffu, emit == 2);
            const u32 off = nn + __popc(b1 & lt) + __popc(b2 & lt);
            if (emit >= 1)
                nxt[off] = o0; // ncur < 64 on entry: at most 126 entries
            if (emit == 2)
                nxt[off + 1] = o1;
            nn += __popc(b1) + __popc(b2);
        }
        u32 *t = cur;
        cur    = nxt;
        nxt    = t;
        ncur   = nn;
        if (!__any_sync(0xffffffffu, pending) || ncur >= TOP_CAP / 2)
            break;
    }
    __syncwarp();
    for (u32 k = lane; k < ncur; k += 32)
        top_front[u64(sg) * TOP_CAP + k] = cur[k];
    if (lane == 0)
        top_count[sg] = ncur;
}

/// The reference walks the tree once per leaf a and keeps the leaves b with
///   hit(a, b) = cella_neigh_b(a, b ⊕ rint_b·R) || cella_neigh_b(a ⊕ rint_a·R, b);
/// internal nodes are pruned with the same test.  Boxes and rint only grow towards the root and every
/// operation of the test is monotonic in floating point, so the set of leaves a walk reaches is exactly
/// {b : hit(a, b)} — however the tree is traversed.  Here ONE warp walks for GL consecutive leaves: the
/// frontier (shared memory, left-to-right order kept by an ordered compaction) is tested 32 nodes at a
/// time, first against the union of the group's boxes (a superset of every member's test: cheap reject),
/// then exactly against each member that still hit the parent; every entry carries the 8-bit mask of
/// the members it hits, so the frontier is the union of the members' own walks and nothing more.
/// Output per group: (first rank, mask << 24 | length) of the candidate leaves in ascending rank order.
/// BIG = false: the frontier lives in shared memory (F entries); a group whose walk does not fit appends
/// itself to `over_list`.  BIG = true: one block per group of `over_list`, frontier in global scratch (F = L
/// entries: a frontier holds roots of disjoint subtrees, never more than there are leaves) — the few groups
/// around a particle with a very large h (the reference searches those too, slowly).
template<bool BIG>
__global__ void __launch_bounds__(WALK_WARPS * 32, 32) group_walk_kernel(
    const NodePack *__restrict__ nodes, u32 I, u32 L, const u32 *__restrict__ real_prefix, f64 Rkern, u32 F, u32 SUPER,
    u32 TOP_CAP, const u32 *__restrict__ top_front, const u32 *__restrict__ top_count, u64 ecap,
    unsigned long long *__restrict__ ecursor, uint2 *__restrict__ gcand, u64 *__restrict__ gc_off,
    u32 *__restrict__ gcount, u32 *__restrict__ flags, u32 *__restrict__ over_list, u32 over_cap,
    uint2 *__restrict__ big_scratch) {
    extern __shared__ __align__(16) unsigned char walk_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = sizeof(LeafBox) * (GL + 1) + (BIG ? 0 : size_t(2) * F * sizeof(uint2));
    LeafBox *lb = reinterpret_cast<LeafBox *>(walk_smem + warp * per_warp); // members, then their union
    uint2 *cur  = BIG ? big_scratch + u64(blockIdx.x) * 2 * F : reinterpret_cast<uint2 *>(lb + GL + 1);
    uint2 *nxt  = cur + F;
    const u32 G  = (L + GL - 1) / GL;
    const u32 g  = BIG ? over_list[blockIdx.x] : blockIdx.x * WALK_WARPS + warp;
    const u32 lt = (1u << lane) - 1u;
    if (g >= G)
        return;
    const f64 inf = __longlong_as_double(0x7ff0000000000000ll);
    f64 l0 = inf, l1 = inf, l2 = inf, h0 = -inf, h1 = -inf, h2 = -inf;
    f64 p0 = inf, p1 = inf, p2 = inf, q0 = -inf, q1 = -inf, q2 = -inf;
    bool valid = false;
    {
        u32 leaf = g * GL + lane;
        if (lane < GL && leaf < L) {
            NodeRegs a = load_node(nodes + I + leaf);
            valid      = real_prefix[a.right] != real_prefix[a.left]; // a leaf of ghosts only needs no list
            f64 ar     = a.rint * Rkern;
            LeafBox b;
            b.lo[0] = a.lo0, b.lo[1] = a.lo1, b.lo[2] = a.lo2, b.hi[0] = a.hi0, b.hi[1] = a.hi1, b.hi[2] = a.hi2;
            b.e0[0] = a.lo0 - ar, b.e0[1] = a.lo1 - ar, b.e0[2] = a.lo2 - ar;
            b.e1[0] = a.hi0 + ar, b.e1[1] = a.hi1 + ar, b.e1[2] = a.hi2 + ar;
            lb[lane] = b;
            if (valid) {
                l0 = b.lo[0], l1 = b.lo[1], l2 = b.lo[2], h0 = b.hi[0], h1 = b.hi[1], h2 = b.hi[2];
                p0 = b.e0[0], p1 = b.e0[1], p2 = b.e0[2], q0 = b.e1[0], q1 = b.e1[1], q2 = b.e1[2];
            }
        }
    }
    const u32 vmask = __ballot_sync(0xffffffffu, valid);
    if (vmask == 0) {
        if (lane == 0)
            gcount[g] = 0;
        return;
    }
    l0 = warp_min8(l0), l1 = warp_min8(l1), l2 = warp_min8(l2);
    h0 = warp_max8(h0), h1 = warp_max8(h1), h2 = warp_max8(h2);
    p0 = warp_min8(p0), p1 = warp_min8(p1), p2 = warp_min8(p2);
    q0 = warp_max8(q0), q1 = warp_max8(q1), q2 = warp_max8(q2);
    // a compact group (the usual case: 8 neighbouring leaves) prunes internal nodes with the union test
    // only; a spread one (Morton-consecutive leaves far apart) runs the member tests at every level
    f64 m0 = -inf, m1 = -inf, m2 = -inf;
    if (valid) {
        const LeafBox &b = lb[lane];
        m0 = b.e1[0] - b.e0[0], m1 = b.e1[1] - b.e0[1], m2 = b.e1[2] - b.e0[2];
    }
    m0 = warp_max8(m0), m1 = warp_max8(m1), m2 = warp_max8(m2);
    const bool spread = (q0 - p0 > 2. * m0) || (q1 - p1 > 2. * m1) || (q2 - p2 > 2. * m2);
    if (lane == 0) {
        LeafBox u; // kept in shared memory (broadcast reads) rather than in 24 registers
        u.lo[0] = l0, u.lo[1] = l1, u.lo[2] = l2, u.hi[0] = h0, u.hi[1] = h1, u.hi[2] = h2;
        u.e0[0] = p0, u.e0[1] = p1, u.e0[2] = p2, u.e1[0] = q0, u.e1[1] = q1, u.e1[2] = q2;
        lb[GL]  = u;
    }
    const LeafBox &U = lb[GL];
    // start: the frontier the super-group's top walk left (untested nodes, left to right); entries are
    // (node, members to test << 24 | 0) — length 0 marks a node that is not tested yet
    const u32 ncur0 = top_count[g / SUPER];
    for (u32 k = lane; k < ncur0; k += 32)
        cur[k] = make_uint2(top_front[u64(g / SUPER) * TOP_CAP + k], vmask << 24);
    u32 ncur  = ncur0;
    bool more = true;
    while (more) {
        __syncwarp();
        u32 nn       = 0;
        bool pending = false;
        for (u32 base = 0; base < ncur; base += 32) {
            const u32 k = base + lane;
            u32 emit    = 0;
            uint2 o0 = make_uint2(0u, 0u), o1 = o0;
            const uint2 e = k < ncur ? cur[k] : make_uint2(0u, 1u);
            if (__all_sync(0xffffffffu, (e.y & 0xffffffu) != 0)) { // only candidate leaves of earlier rounds: copy
                const u32 cnt = min(32u, ncur - base);
                if (nn + cnt <= F && k < ncur)
                    nxt[nn + lane] = e;
                nn += cnt;
                continue;
            }
            if (k < ncur) {
                if (e.y & 0xffffffu) { // a candidate leaf found in an earlier round
                    emit = 1;
                    o0   = e;
                } else { // node e.x, to be tested for the members in e.y >> 24 (those that hit its parent)
                    NodeRegs n = load_node(nodes + e.x);
                    f64 r      = n.rint * Rkern;
                    f64 x1 = n.hi0 + r, y1 = n.hi1 + r, z1 = n.hi2 + r;
                    f64 x0 = n.lo0 - r, y0 = n.lo1 - r, z0 = n.lo2 - r;
                    // fmax(x,y) <= fmin(u,v)  <=>  x <= v && y <= u when x <= u and y <= v hold by construction
                    bool hit = ((U.lo[0] <= x1) & (x0 <= U.hi[0]) & (U.lo[1] <= y1) & (y0 <= U.hi[1])
                                & (U.lo[2] <= z1) & (z0 <= U.hi[2]))
                               | ((U.e0[0] <= n.hi0) & (n.lo0 <= U.e1[0]) & (U.e0[1] <= n.hi1) & (n.lo1 <= U.e1[1])
                                  & (U.e0[2] <= n.hi2) & (n.lo2 <= U.e1[2]));
                    u32 mask = 0;
                    if (hit) {
                        const u32 pm = e.y >> 24;
                        if (e.x >= I || spread) { // the exact test of the reference, member by member
#pragma unroll
                            for (int i = 0; i < GL; i++) {
                                if (!((vmask >> i) & 1u)) // warp-uniform
                                    continue;
                                const LeafBox &A = lb[i];
                                bool c1 = (A.lo[0] <= x1) & (x0 <= A.hi[0]) & (A.lo[1] <= y1) & (y0 <= A.hi[1])
                                          & (A.lo[2] <= z1) & (z0 <= A.hi[2]);
                                bool c2 = (A.e0[0] <= n.hi0) & (n.lo0 <= A.e1[0]) & (A.e0[1] <= n.hi1)
                                          & (n.lo1 <= A.e1[1]) & (A.e0[2] <= n.hi2) & (n.lo2 <= A.e1[2]);
                                if (c1 | c2)
                                    mask |= 1u << i;
                            }
                            mask &= pm;
                        } else {
                            mask = pm;
                        }
                    }
                    if (mask) {
                        if (e.x >= I) { // leaf
                            u32 len = n.right - n.left;
                            if (len >= (1u << 24))
                                atomicOr(flags + 2, 1u);
                            emit = 1;
                            o0   = make_uint2(e.x, (mask << 24) | (len & 0xffffffu)); // node id: see euclid_cull_kernel
                        } else {
                            emit    = 2;
                            o0      = make_uint2(n.left, mask << 24);
                            o1      = make_uint2(n.right, mask << 24);
                            pending = true;
                        }
                    }
                }
            }
            // ordered compaction: emit is 0, 1 or 2 -> two ballots give the exclusive prefix
            const u32 b1 = __ballot_sync(0xffffffffu, emit >= 1), b2 = __ballot_sync(0xffffffffu, emit == 2);
            const u32 total = __popc(b1) + __popc(b2);
            const u32 off   = nn + __popc(b1 & lt) + __popc(b2 & lt);
            if (nn + total <= F) {
                if (emit >= 1)
                    nxt[off] = o0;
                if (emit == 2)
                    nxt[off + 1] = o1;
            }
            nn += total;
        }
        more = __any_sync(0xffffffffu, pending);
        if (nn > F) { // the frontier does not fit: this group is walked again with a frontier in global memory
            if (lane == 0) {
                u32 slot = atomicAdd(flags + 0, 1u);
                if (slot < over_cap && !BIG)
                    over_list[slot] = g;
                if (BIG)
                    atomicOr(flags + 2, 2u);
                gcount[g] = 0;
            }
            return;
        }
        uint2 *t = cur;
        cur      = nxt;
        nxt      = t;
        ncur     = nn;
    }
    __syncwarp();
    // output: a slice of the candidate-entry array reserved with one atomic
    unsigned long long ebase = 0;
    if (lane == 0)
        ebase = atomicAdd(ecursor, (unsigned long long) ncur);
    ebase = __shfl_sync(0xffffffffu, ebase, 0);
    if (ebase + ncur > ecap) { // the host reads the cursor, grows the array and repeats the search
        if (lane == 0)
            gcount[g] = 0;
        return;
    }
    for (u32 k = lane; k < ncur; k += 32)
        gcand[ebase + k] = cur[k];
    if (lane == 0) {
        gc_off[g] = ebase;
        gcount[g] = ncur;
    }
}

// ---------------------------------------------------------------------------------------------
// stage 1b: Euclidean cull of the candidate leaves
// ---------------------------------------------------------------------------------------------
/// The reference's leaf test compares boxes axis by axis: a candidate leaf b passes for leaf A when its box
/// reaches into A's box grown by R on every axis — a CUBE around A.  A pair (a in A, b' in b) is only accepted
/// when r <= R_a or r <= R_b', and r is at least the Euclidean distance of the two boxes, so a candidate farther
/// than max(R_A, R_b) from A (the corners of the cube: a quarter of the candidates on a lattice) cannot
/// contribute to the lists of A.  One warp per group clears those member bits (1e-9 safety margin on the
/// rounding of the distance: the cull is conservative, the lists are unchanged) and replaces the node id the
/// walk left in the entry by the leaf's first rank.  Lanes = entries: the work is two node loads and
/// eight box distances per entry, fully parallel — the same test inside the walk costs 4 ms, here 1.
__global__ void __launch_bounds__(128) euclid_cull_kernel(
    const NodePack *__restrict__ nodes, u32 I, u32 L, f64 Rkern, uint2 *__restrict__ gcand,
    const u64 *__restrict__ gc_off, const u32 *__restrict__ gcount, const u32 *__restrict__ group_ids, u32 ngroups) {
    __shared__ f64 mb[4][GL][8]; // per warp: the members' boxes (lo, hi) and interaction radius
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 gi   = blockIdx.x * 4 + warp;
    if (gi >= ngroups)
        return;
    const u32 g  = group_ids ? group_ids[gi] : gi;
    const u32 gn = gcount[g];
    if (gn == 0)
        return;
    if (lane < GL) {
        const u32 leaf = g * GL + lane;
        f64 *m         = mb[warp][lane];
        if (leaf < L) {
            NodeRegs a = load_node(nodes + I + leaf);
            m[0] = a.lo0, m[1] = a.lo1, m[2] = a.lo2, m[3] = a.hi0, m[4] = a.hi1, m[5] = a.hi2, m[6] = a.rint * Rkern;
        } else {
            m[0] = m[1] = m[2] = m[3] = m[4] = m[5] = m[6] = 0;
        }
    }
    __syncwarp();
    uint2 *gc = gcand + gc_off[g];
    for (u32 k = lane; k < gn; k += 32) {
        uint2 e    = gc[k];
        NodeRegs n = load_node(nodes + e.x);
        const f64 r = n.rint * Rkern;
        u32 mask    = e.y >> 24;
#pragma unroll
        for (int i = 0; i < GL; i++) {
            const f64 *m = mb[warp][i];
            f64 g0 = fmax(fmax(m[0] - n.hi0, n.lo0 - m[3]), 0.);
            f64 g1 = fmax(fmax(m[1] - n.hi1, n.lo1 - m[4]), 0.);
            f64 g2 = fmax(fmax(m[2] - n.hi2, n.lo2 - m[5]), 0.);
            f64 Rm = fmax(m[6], r);
            if (g0 * g0 + g1 * g1 + g2 * g2 > Rm * Rm * (1. + 1e-9))
                mask &= ~(1u << i);
        }
        gc[k] = make_uint2(n.left, (mask << 24) | (e.y & 0xffffffu));
    }
}

// ---------------------------------------------------------------------------------------------
// stage 2: accept pass (ballots) and ordered fill, warp per leaf, block per group
// ---------------------------------------------------------------------------------------------
constexpr u32 RANK_CACHE   = 1024; ///< per-warp window of candidate ranks
constexpr u32 BALLOT_WORDS = 512;  ///< per-warp ballot store: (particles of the batch) x (chunks)

struct alignas(32) WarpScratch {
    Pack4 pa[32];          // (x, y, z, lim_a) of the current batch of particles
    u32 rankc[RANK_CACHE]; // candidate ranks of the current window
    u32 ball[BALLOT_WORDS];
};

/// per-batch state of one warp while it scans the candidates of its leaf
struct ScanState {
    u32 nb;      ///< particles in the batch
    u32 cglob;   ///< chunks processed so far
    u32 mycount; ///< lane a < nb: accepted so far (MODE 0) / write cursor (MODE 1)
    bool fit;    ///< every ballot so far is in the shared store
    u32 ncand;   ///< candidates of the leaf seen so far
};

/// NC chunks of 32 candidates (lane = one candidate of each chunk) against the nb particles of the batch;
/// two chunks per pass halve the shared-memory reads of the particles and give two independent chains.
/// MODE 0: ballots + counts;  MODE 1: test again and write (used when the ballots / ranks did not fit)
/// V = 1 (default): a candidate slot past the end gets coordinates no particle is near to (r² = inf, so
/// its ballot bit is clear without a validity term), and the per-particle counts are taken from the kept
/// ballots after the particle loop instead of inside it.  V = 0: the first version, kept for tuning runs.
template<int MODE, int NC, int V>
__device__ __forceinline__ void chunk_body(
    WarpScratch &w, ScanState &st, const Pack4 *__restrict__ SA, const bool (&vb)[NC], const u32 (&rank_b)[NC],
    f64 Rker2, f64 h_tolerance, int lane, u32 lt, u32 *__restrict__ list_s) {
    f64 bx[NC], by[NC], bz[NC], lim_b[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        bx[c] = V ? 1e300 : 0.;
        by[c] = bz[c] = lim_b[c] = 0;
        if (vb[c]) {
            Pack4 q = ld4(SA + rank_b[c]);
            bx[c] = q.a, by[c] = q.b, bz[c] = q.c;
            f64 rint_b = q.d * h_tolerance;
            lim_b[c]   = rint_b * rint_b * Rker2;
        }
    }
    const bool keep = MODE == 0 && (st.cglob + NC) * st.nb <= BALLOT_WORDS;
    u32 mymask[NC];
#pragma unroll
    for (int c = 0; c < NC; c++)
        mymask[c] = 0;
    for (u32 a = 0; a < st.nb; a++) {
        Pack4 pa = w.pa[a];
        u32 m[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) {
            f64 dx = pa.a - bx[c], dy = pa.b - by[c], dz = pa.c - bz[c];
            f64 rab2         = dx * dx + dy * dy + dz * dz;
            bool no_interact = rab2 > pa.d && rab2 > lim_b[c];
            m[c]             = __ballot_sync(0xffffffffu, V ? !no_interact : (vb[c] && !no_interact));
        }
        if (MODE == 0) {
            if (lane == int(a)) {
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    mymask[c] = m[c];
                    if (!V)
                        st.mycount += __popc(m[c]);
                }
            }
        } else {
            u32 bo = __shfl_sync(0xffffffffu, st.mycount, a);
#pragma unroll
            for (int c = 0; c < NC; c++) {
                if ((m[c] >> lane) & 1u)
                    list_s[bo + __popc(m[c] & lt)] = rank_b[c];
                bo += __popc(m[c]);
            }
            if (lane == int(a))
                st.mycount = bo;
        }
    }
    if (MODE == 0) {
        if (V) { // mymask is zero in the lanes that hold no particle
#pragma unroll
            for (int c = 0; c < NC; c++)
                st.mycount += __popc(mymask[c]);
        }
        if (keep) {
            if (lane < int(st.nb)) {
#pragma unroll
                for (int c = 0; c < NC; c++)
                    w.ball[(st.cglob + c) * st.nb + lane] = mymask[c];
            }
        } else {
            st.fit = false;
        }
    }
    st.cglob += NC;
}

/// the window of `run` ranks in w.rankc, 64 (then 32) candidates at a time
template<int MODE, int V>
__device__ __forceinline__ void flush_window(
    WarpScratch &w, ScanState &st, const Pack4 *__restrict__ SA, u32 run, f64 Rker2, f64 h_tolerance, int lane, u32 lt,
    u32 *__restrict__ list_s) {
    __syncwarp();
    u32 j0 = 0;
    for (; j0 + 32 < run; j0 += 64) {
        const u32 jA = j0 + lane, jB = j0 + 32 + lane;
        const bool vb[2]  = {true, jB < run};
        const u32 rk[2]   = {w.rankc[jA], vb[1] ? w.rankc[jB] : 0u};
        chunk_body<MODE, 2, V>(w, st, SA, vb, rk, Rker2, h_tolerance, lane, lt, list_s);
    }
    if (j0 < run) {
        const u32 j      = j0 + lane;
        const bool vb[1] = {j < run};
        const u32 rk[1]  = {vb[0] ? w.rankc[j] : 0u};
        chunk_body<MODE, 1, V>(w, st, SA, vb, rk, Rker2, h_tolerance, lane, lt, list_s);
    }
    __syncwarp();
}

/// scans the candidates of one leaf of group g: the group's entries that carry the leaf's bit, expanded
/// into windows of ranks (shared memory) and processed 32 at a time.  Returns true when everything went
/// through ONE window that is still in w.rankc (then `nwin` is its length and the ballots can be replayed).
template<int MODE, int V>
__device__ __forceinline__ bool scan_candidates(
    WarpScratch &w, ScanState &st, const uint2 *__restrict__ gc, u32 gn, u32 bit, const Pack4 *__restrict__ SA,
    f64 Rker2, f64 h_tolerance, int lane, u32 lt, u32 *__restrict__ list_s, u32 &nwin) {
    u32 run     = 0;
    bool single = true;
    for (u32 base = 0; base < gn; base += 32) {
        u32 k   = base + lane;
        uint2 e = k < gn ? gc[k] : make_uint2(0u, 0u);
        u32 len = (e.y & bit) ? (e.y & 0xffffffu) : 0u;
        u32 inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o)
                inc += t;
        }
        const u32 tot = __shfl_sync(0xffffffffu, inc, 31);
        if (tot == 0)
            continue;
        st.ncand += tot;
        if (tot > RANK_CACHE) { // very large leaves (many equal Morton codes): straight from the entries
            if (run)
                flush_window<MODE, V>(w, st, SA, run, Rker2, h_tolerance, lane, lt, list_s);
            run    = 0;
            single = false;
            for (int kk = 0; kk < 32; kk++) {
                u32 s0 = __shfl_sync(0xffffffffu, e.x, kk), ln = __shfl_sync(0xffffffffu, len, kk);
                for (u32 j0 = 0; j0 < ln; j0 += 32) {
                    const u32 j      = j0 + lane;
                    const bool vb[1] = {j < ln};
                    const u32 rk[1]  = {s0 + j};
                    chunk_body<MODE, 1, V>(w, st, SA, vb, rk, Rker2, h_tolerance, lane, lt, list_s);
                }
            }
            continue;
        }
        if (run + tot > RANK_CACHE) {
            flush_window<MODE, V>(w, st, SA, run, Rker2, h_tolerance, lane, lt, list_s);
            run    = 0;
            single = false;
        }
        u32 dst = run + inc - len;
        if (V) { // the longest run of the 32 entries bounds the (warp-uniform) trip count
            const u32 maxlen = __reduce_max_sync(0xffffffffu, len);
            u32 *wr = w.rankc + dst;
            for (u32 t = 0; t < maxlen; t++)
                if (t < len)
                    wr[t] = e.x + t;
        } else {
            for (u32 t = 0; __any_sync(0xffffffffu, t < len); t++)
                if (t < len)
                    w.rankc[dst + t] = e.x + t;
        }
        run += tot;
    }
    nwin = run;
    if (run)
        flush_window<MODE, V>(w, st, SA, run, Rker2, h_tolerance, lane, lt, list_s);
    return single;
}

constexpr int S2_WARPS = GL; // one block = one group of leaves

/// Stage 2, one launch: warp per leaf.  Pass 1 tests every (particle, candidate) pair and keeps the ballots
/// in shared memory; the warp then reserves the leaf's list space with one atomic on a global cursor
/// (lists are contiguous per leaf, leaves in completion order — the internal CSR does not need a global
/// order, the exported ObjectCache is rebuilt by id) and pass 2 replays the ballots.  Leaves whose
/// candidates or ballots do not fit the shared stores test again in pass 2 instead.
template<int V>
__global__ void __launch_bounds__(S2_WARPS * 32, 4) neigh_lists_kernel(
    const NodePack *__restrict__ nodes, u32 I, u32 L, const Pack4 *__restrict__ SA, const u8 *__restrict__ real_flag,
    const u32 *__restrict__ real_prefix, const uint2 *__restrict__ gcand, const u64 *__restrict__ gc_off,
    const u32 *__restrict__ gcount, const u32 *__restrict__ group_ids, f64 Rker2, f64 h_tolerance, u64 list_cap,
    unsigned long long *__restrict__ cursor, u32 *__restrict__ cnt_s, u32 *__restrict__ off_s,
    u32 *__restrict__ list_s) {
    extern __shared__ __align__(32) unsigned char s2_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 lt   = (1u << lane) - 1u;
    const u32 g    = group_ids ? group_ids[blockIdx.x] : blockIdx.x;
    const u32 leaf = g * GL + warp;
    if (leaf >= L)
        return;
    const u32 gn = gcount[g];
    if (gn == 0)
        return;
    WarpScratch &w = reinterpret_cast<WarpScratch *>(s2_smem)[warp];
    const u32 p0   = reinterpret_cast<const u32 *>(&nodes[I + leaf].rint)[2];
    const u32 p1   = reinterpret_cast<const u32 *>(&nodes[I + leaf].rint)[3];
    const u32 slot0 = real_prefix[p0];
    if (real_prefix[p1] == slot0)
        return; // no real particle in this leaf
    const uint2 *gc = gcand + gc_off[g];
    const u32 bit   = 1u << (24 + warp);
    u32 a_base      = 0;
    for (u32 rb = p0; rb < p1; rb += 32) {
        u32 r   = rb + lane;
        bool va = r < p1 && real_flag[r];
        u32 bal = __ballot_sync(0xffffffffu, va);
        u32 nb  = __popc(bal);
        if (nb == 0)
            continue;
        __syncwarp();
        if (va) {
            Pack4 q    = ld4(SA + r);
            f64 rint_a = q.d * h_tolerance;
            w.pa[__popc(bal & lt)] = Pack4{q.a, q.b, q.c, rint_a * rint_a * Rker2};
        }
        __syncwarp();
        // ---- pass 1: ballots + counts
        ScanState st{nb, 0u, 0u, true, 0u};
        u32 nwin    = 0;
        bool single = scan_candidates<0, V>(w, st, gc, gn, bit, SA, Rker2, h_tolerance, lane, lt, list_s, nwin);
        const u32 mycount = lane < int(nb) ? st.mycount : 0u;
        // ---- reserve the list space of this batch: exclusive prefix of the counts + one atomic
        u32 inc = mycount;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o)
                inc += t;
        }
        u32 total = __shfl_sync(0xffffffffu, inc, 31);
        unsigned long long base = 0;
        if (lane == 0) {
            base = atomicAdd(cursor, (unsigned long long) total);
            atomicAdd(cursor + 1, (unsigned long long) nb * st.ncand); // pair tests done (statistics)
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        u32 myoff = u32(base) + inc - mycount;
        if (lane < int(nb)) {
            cnt_s[slot0 + a_base + lane] = mycount;
            off_s[slot0 + a_base + lane] = myoff;
        }
        a_base += nb;
        if (base + total > list_cap)
            continue; // the list array is too small: the host reads the cursor and runs the search again
        __syncwarp();
        // ---- pass 2: ordered fill
        if (single && st.fit) { // replay the ballots of the one window
            const u32 nchunk = (nwin + 31) >> 5;
            for (u32 c = 0; c < nchunk; c += 2) { // two chunks per pass
                const u32 jA = c * 32 + lane, jB = jA + 32;
                const bool two = c + 1 < nchunk;
                const u32 rkA = jA < nwin ? w.rankc[jA] : 0u, rkB = jB < nwin ? w.rankc[jB] : 0u;
                const u32 mkA = lane < int(nb) ? w.ball[c * nb + lane] : 0u;
                const u32 mkB = two && lane < int(nb) ? w.ball[(c + 1) * nb + lane] : 0u;
                for (u32 a = 0; a < nb; a++) {
                    u32 mA = __shfl_sync(0xffffffffu, mkA, a), mB = __shfl_sync(0xffffffffu, mkB, a);
                    u32 bo = __shfl_sync(0xffffffffu, myoff, a);
                    if ((mA >> lane) & 1u)
                        list_s[bo + __popc(mA & lt)] = rkA;
                    if ((mB >> lane) & 1u)
                        list_s[bo + __popc(mA) + __popc(mB & lt)] = rkB;
                }
                myoff += __popc(mkA) + __popc(mkB);
            }
        } else { // did not fit: scan and test again, writing as we go
            ScanState st2{nb, 0u, myoff, true, 0u};
            scan_candidates<1, V>(w, st2, gc, gn, bit, SA, Rker2, h_tolerance, lane, lt, list_s, nwin);
        }
    }
}

// The search of one patch in three pieces, so that a model with many patches can enqueue the searches of all
// of them and synchronise ONCE (Model::start_neighbors_cache): setup (packed nodes, buffers), one attempt
// (walk + cull + lists + the read-back of the cursors, no synchronisation), and the check of an attempt after the
// stream has been synchronised (true: a capacity was exceeded, the exact need is known, run another attempt).
namespace {
constexpr u32 OVER_CAP = 1u << 16;
using ListsKernel = void (*)(
    const NodePack *, u32, u32, const Pack4 *, const u8 *, const u32 *, const uint2 *, const u64 *, const u32 *,
    const u32 *, f64, f64, u64, unsigned long long *, u32 *, u32 *, u32 *);
ListsKernel lists_kernel_choice() { // SHAMB200_NL = 0 selects the first version of the list kernel (tuning runs)
    const char *nl_env = getenv("SHAMB200_NL");
    return (nl_env && atoi(nl_env) == 0) ? neigh_lists_kernel<0> : neigh_lists_kernel<1>;
}
u32 *search_flags(SearchBuffers &sb) { return reinterpret_cast<u32 *>(sb.scalars.p + 2); } // [0] groups over the frontier, [2] errors
unsigned long long *search_cursor(SearchBuffers &sb) { return reinterpret_cast<unsigned long long *>(sb.scalars.p + 4); }
unsigned long long *search_ecursor(SearchBuffers &sb) { return reinterpret_cast<unsigned long long *>(sb.scalars.p + 6); }
constexpr size_t s2_bytes = sizeof(WarpScratch) * S2_WARPS;

void search_setup(cudaStream_t s, const TreeBuffers &tb, SearchBuffers &sb, const f64 *d_rint) {
    const u32 I = tb.I, L = tb.L;
    sb.nodes.ensure(size_t(I) + L, 1.1);
    pack_nodes_kernel<<<grid_for(size_t(I) + L, 256), 256, 0, s>>>(
        I, L, tb.aabb_min.p, tb.aabb_max.p, d_rint, tb.lchild.p, tb.rchild.p, tb.lflag.p, tb.rflag.p,
        tb.reduc_index_map.p, sb.nodes.p);
    SB_COUNT_LAUNCH();
    const u32 G = (L + GL - 1) / GL;
    sb.gcount.ensure(G, 1.1);
    sb.gc_off.ensure(G, 1.1);
    sb.cnt_s.ensure(sb.N, 1.1);
    sb.off_s.ensure(sb.N, 1.1);
    static bool attr_set = false;
    if (!attr_set) {
        SB_CUDA_CHECK(cudaFuncSetAttribute(neigh_lists_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(s2_bytes)));
        SB_CUDA_CHECK(cudaFuncSetAttribute(neigh_lists_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(s2_bytes)));
        SB_CUDA_CHECK(cudaFuncSetAttribute(
            group_walk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    // capacities: what the previous search of this patch needed (+ slack); the first time ~96 list entries
    // per particle and ~24 candidate entries per leaf.  ONE host synchronisation per search: it reads the
    // list cursor (= K), the candidate-entry cursor and the number of groups whose walk did not fit its
    // shared-memory frontier; if a capacity was exceeded the exact need is known and the search is repeated.
    if (sb.list_s.cap == 0)
        sb.list_s.ensure(size_t(sb.N) * 96 + 1024);
    if (sb.gcand.cap == 0)
        sb.gcand.ensure(size_t(L) * 24 + 4096);
    // head room: a search that outgrows an array runs twice.  Lists grow from step to step while a system
    // relaxes (h grows by several per cent per step in a fresh disc), so an array the last search filled to more
    // than 85 % is replaced before this one starts (the old lists are not needed any more).
    if (sb.K && double(sb.K) > 0.85 * double(sb.list_s.cap))
        sb.list_s.ensure(size_t(double(sb.K) * 1.4));
    if (sb.entries_last && double(sb.entries_last) > 0.85 * double(sb.gcand.cap))
        sb.gcand.ensure(size_t(double(sb.entries_last) * 1.4));
    sb.over_list.ensure(OVER_CAP);
    // tuning knobs (scripts/tune.py): shared-memory frontier entries per walk and the walk kernel's shared-memory
    // carve-out (per cent of the SM's unified L1 / shared array); both bound the resident walks per SM
    if (const char *e = getenv("SHAMB200_WALK_F"))
        sb.frontier_cap = std::max(64, atoi(e));
    if (const char *e = getenv("SHAMB200_WALK_CARVE"))
        SB_CUDA_CHECK(cudaFuncSetAttribute(
            group_walk_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, std::min(100, std::max(0, atoi(e)))));
}

/// walk + cull + lists of every group and the read-back of the cursors; no synchronisation
void search_attempt(
    cudaStream_t s, const TreeBuffers &tb, SearchBuffers &sb, f64 Rkern, f64 h_tolerance,
    const std::function<void(const char *)> &mark) {
    const u32 I = tb.I, L = tb.L, G = (L + GL - 1) / GL;
    const f64 Rker2       = Rkern * Rkern;
    const size_t per_warp = sizeof(LeafBox) * (GL + 1) + size_t(2) * sb.frontier_cap * sizeof(uint2);
    const u64 ecap        = sb.gcand.cap;
    SB_CUDA_CHECK(cudaMemsetAsync(sb.scalars.p + 2, 0, 5 * sizeof(u64), s));
    if (mark)
        mark("neigh_walk");
    const u32 SUPER = top_cfg().super, TOP_CAP = top_cfg().cap;
    const u32 S = (G + SUPER - 1) / SUPER;
    sb.top_front.ensure(size_t(S) * TOP_CAP, 1.1);
    sb.top_count.ensure(S, 1.1);
    top_walk_kernel<<<S, 32, 0, s>>>(sb.nodes.p, I, L, sb.real_prefix.p, Rkern, SUPER, TOP_CAP, sb.top_front.p, sb.top_count.p);
    SB_COUNT_LAUNCH();
    group_walk_kernel<false><<<grid_for(G, WALK_WARPS), WALK_WARPS * 32, per_warp * WALK_WARPS, s>>>(
        sb.nodes.p, I, L, sb.real_prefix.p, Rkern, sb.frontier_cap, SUPER, TOP_CAP, sb.top_front.p, sb.top_count.p, ecap,
        search_ecursor(sb), sb.gcand.p, sb.gc_off.p, sb.gcount.p, search_flags(sb), sb.over_list.p, OVER_CAP, nullptr);
    SB_COUNT_LAUNCH();
    euclid_cull_kernel<<<grid_for(G, 4), 128, 0, s>>>(
        sb.nodes.p, I, L, Rkern, sb.gcand.p, sb.gc_off.p, sb.gcount.p, nullptr, G);
    SB_COUNT_LAUNCH();
    if (mark)
        mark("neigh_lists");
    lists_kernel_choice()<<<G, S2_WARPS * 32, s2_bytes, s>>>(
        sb.nodes.p, I, L, sb.SA.p, sb.real_flag.p, sb.real_prefix.p, sb.gcand.p, sb.gc_off.p, sb.gcount.p, nullptr,
        Rker2, h_tolerance, u64(sb.list_s.cap), search_cursor(sb), sb.cnt_s.p, sb.off_s.p, sb.list_s.p);
    SB_COUNT_LAUNCH();
    d2h_small(s, sb.h_scalars.p + 2, sb.scalars.p + 2, 5 * sizeof(u64));
}

void search_errors(const SearchBuffers &sb) {
    const u32 err = u32(sb.h_scalars.p[3] & 0xffffffffull);
    if (err & 1u)
        throw std::runtime_error("neighbour search: a tree leaf holds 2^24 or more objects");
    if (err & 2u)
        throw std::runtime_error("neighbour search: internal error, a global-memory frontier overflowed");
}

/// after the synchronisation that follows an attempt: true = run another attempt (capacities adjusted)
bool search_check(
    cudaStream_t s, const TreeBuffers &tb, SearchBuffers &sb, f64 Rkern, f64 h_tolerance, int attempt) {
    const u32 I = tb.I, L = tb.L, G = (L + GL - 1) / GL;
    const f64 Rker2 = Rkern * Rkern;
    const u64 ecap  = sb.gcand.cap;
    search_errors(sb);
    u32 n_over = u32(sb.h_scalars.p[2] & 0xffffffffull);
    bool redo  = false;
    // many groups over the frontier: a larger shared-memory frontier for everybody; a few (the
    // surroundings of a particle with a very large h): those groups again with a frontier in global memory
    if (n_over > OVER_CAP || (n_over > G / 64 && sb.frontier_cap < 2048)) {
        if (sb.frontier_cap >= 8192)
            throw std::runtime_error("neighbour search: too many leaf groups need a very large tree-walk frontier");
        // x 1.4 in multiples of 64 entries (320, 448, 640, 896, ...): the frontier is what bounds the resident
        // walks per SM, and doubling overshoots (17 M particles: 448 entries suffice, walk 5.5 -> 5.1 ms)
        sb.frontier_cap = (sb.frontier_cap * 7 / 5 + 63) / 64 * 64;
        redo = true;
    } else if (n_over > 0 && sb.h_scalars.p[6] <= ecap) {
        const u64 budget = 4ull << 30; // scratch for the global frontiers, groups in batches
        const u32 batch  = u32(std::max<u64>(1, std::min<u64>(n_over, budget / (u64(2) * L * sizeof(uint2)))));
        sb.big_scratch.ensure(size_t(batch) * 2 * L);
        for (u32 b0 = 0; b0 < n_over; b0 += batch) {
            const u32 nb = std::min(batch, n_over - b0);
            group_walk_kernel<true><<<nb, WALK_WARPS * 32, sizeof(LeafBox) * (GL + 1), s>>>(
                sb.nodes.p, I, L, sb.real_prefix.p, Rkern, L, top_cfg().super, top_cfg().cap, sb.top_front.p, sb.top_count.p, ecap,
                search_ecursor(sb), sb.gcand.p, sb.gc_off.p, sb.gcount.p, search_flags(sb), sb.over_list.p + b0, OVER_CAP,
                sb.big_scratch.p);
            SB_COUNT_LAUNCH();
            euclid_cull_kernel<<<grid_for(nb, 4), 128, 0, s>>>(
                sb.nodes.p, I, L, Rkern, sb.gcand.p, sb.gc_off.p, sb.gcount.p, sb.over_list.p + b0, nb);
            SB_COUNT_LAUNCH();
            lists_kernel_choice()<<<nb, S2_WARPS * 32, s2_bytes, s>>>(
                sb.nodes.p, I, L, sb.SA.p, sb.real_flag.p, sb.real_prefix.p, sb.gcand.p, sb.gc_off.p, sb.gcount.p,
                sb.over_list.p + b0, Rker2, h_tolerance, u64(sb.list_s.cap), search_cursor(sb), sb.cnt_s.p, sb.off_s.p,
                sb.list_s.p);
            SB_COUNT_LAUNCH();
        }
        d2h_small(s, sb.h_scalars.p + 2, sb.scalars.p + 2, 5 * sizeof(u64));
        SB_CUDA_CHECK(cudaStreamSynchronize(s));
        search_errors(sb);
    }
    const u64 need_e = sb.h_scalars.p[6];
    sb.entries_last  = need_e;
    sb.K             = sb.h_scalars.p[4];
    sb.pair_tests    = sb.h_scalars.p[5];
    if (need_e > ecap) { // the candidate-entry array is too small
        sb.gcand.ensure(need_e, 1.25);
        redo = true;
    }
    if (!redo && sb.K > 0xFFFFFFFFull)
        throw std::overflow_error(
            "neighbour count overflows u32 (sum_neigh_cnt is u32 in the reference, TreeTraversal.hpp:378): "
            "use more / smaller patches");
    if (!redo && sb.K > sb.list_s.cap) {
        sb.list_s.ensure(sb.K, 1.25); // lists grow while a system relaxes: room for a few steps, not for one
        redo = true;
    }
    sb.attempts_last    = u32(attempt) + 1;
    sb.over_groups_last = n_over;
    if (redo && attempt >= 8)
        throw std::runtime_error("neighbour search: capacity retry failed");
    return redo;
}
} // namespace

void search_build(
    cudaStream_t s, const TreeBuffers &tb, SearchBuffers &sb, const f64 *d_rint, f64 Rkern, f64 h_tolerance,
    const std::function<void(const char *)> &mark) {
    search_setup(s, tb, sb, d_rint);
    for (int attempt = 0;; attempt++) {
        search_attempt(s, tb, sb, Rkern, h_tolerance, mark);
        SB_CUDA_CHECK(cudaStreamSynchronize(s));
        if (!search_check(s, tb, sb, Rkern, h_tolerance, attempt))
            break;
    }
    SB_LAUNCH_CHECK();
}
void search_enqueue(
    cudaStream_t s, const TreeBuffers &tb, SearchBuffers &sb, const f64 *d_rint, f64 Rkern, f64 h_tolerance) {
    search_setup(s, tb, sb, d_rint);
    search_attempt(s, tb, sb, Rkern, h_tolerance, nullptr);
}
void search_finish(cudaStream_t s, const TreeBuffers &tb, SearchBuffers &sb, f64 Rkern, f64 h_tolerance) {
    for (int attempt = 0; search_check(s, tb, sb, Rkern, h_tolerance, attempt); attempt++) {
        search_attempt(s, tb, sb, Rkern, h_tolerance, nullptr);
        SB_CUDA_CHECK(cudaStreamSynchronize(s));
    }
    SB_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------
// export: ObjectCache of the reference (cnt_neigh / scanned_cnt by id, index_neigh_map of ids)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) export_cnt_kernel(
    const u32 *__restrict__ cnt_s, const u32 *__restrict__ slot_rank, const u32 *__restrict__ index_map, u32 N,
    u32 *__restrict__ x_cnt) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < N)
        x_cnt[index_map[slot_rank[k]]] = cnt_s[k];
}
/// 8 lanes per particle copy its list, ranks -> ids
__global__ void __launch_bounds__(256) export_list_kernel(
    const u32 *__restrict__ cnt_s, const u32 *__restrict__ off_s, const u32 *__restrict__ list_s,
    const u32 *__restrict__ slot_rank, const u32 *__restrict__ index_map, u32 N, const u32 *__restrict__ x_scanned,
    u32 *__restrict__ x_list) {
    u64 t  = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    u32 k  = u32(t >> 3);
    u32 sl = u32(t & 7);
    if (k >= N)
        return;
    u32 id  = index_map[slot_rank[k]];
    u32 c   = cnt_s[k];
    u32 src = off_s[k], dst = x_scanned[id];
    for (u32 j = sl; j < c; j += 8)
        x_list[dst + j] = index_map[list_s[src + j]];
}

void export_object_cache(cudaStream_t s, const TreeBuffers &tb, SearchBuffers &sb) {
    if (sb.exported)
        return;
    sb.x_cnt.ensure(sb.N, 1.1);
    sb.x_scanned.ensure(sb.N, 1.1);
    sb.x_list.ensure(sb.K, 1.05);
    export_cnt_kernel<<<grid_for(sb.N, 256), 256, 0, s>>>(sb.cnt_s.p, sb.slot_rank.p, tb.index_map.p, sb.N, sb.x_cnt.p);
    SB_COUNT_LAUNCH();
    exclusive_scan<u32>(s, sb.x_cnt.p, sb.x_scanned.p, sb.N, sb.scan_tmp, sb.scalars.p + 5);
    export_list_kernel<<<grid_for(u64(sb.N) * 8, 256), 256, 0, s>>>(
        sb.cnt_s.p, sb.off_s.p, sb.list_s.p, sb.slot_rank.p, tb.index_map.p, sb.N, sb.x_scanned.p, sb.x_list.p);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
    SB_CUDA_CHECK(cudaStreamSynchronize(s));
    sb.exported = true;
}

} // namespace sb
