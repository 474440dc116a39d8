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
// once per leaf and emits merged candidate rank ranges into a capped slot (no count pass).  Stage 2 is
// warp-per-leaf with the 32 lanes holding 32 CANDIDATES; the leaf's particles are broadcast from shared
// memory, each (particle, chunk) produces one ballot word.  The fill pass replays the ballots only.
#include "neigh2.cuh"

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
// stage 1: candidate rank ranges of every leaf (one capped walk, no count pass)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool cella_neigh_b2(
    f64 ax0, f64 ay0, f64 az0, f64 ax1, f64 ay1, f64 az1, f64 bx0, f64 by0, f64 bz0, f64 bx1, f64 by1, f64 bz1) {
    return (fmax(ax0, bx0) <= fmin(ax1, bx1)) && (fmax(ay0, by0) <= fmin(ay1, by1))
           && (fmax(az0, bz0) <= fmin(az1, bz1));
}

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

constexpr int WALK_STACK = 32; // tree depth <= 30 (u32 Morton codes) + 1

__global__ void __launch_bounds__(128) leaf_ranges_kernel(
    const NodePack *__restrict__ nodes, u32 I, u32 L, const u32 *__restrict__ real_prefix, f64 Rkern, u32 cap,
    u32 *__restrict__ ranges, u32 *__restrict__ nrange, u32 *__restrict__ ncand, u32 *__restrict__ max_ranges) {
    u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= L)
        return;
    NodeRegs a = load_node(nodes + I + g);
    u32 nreal  = real_prefix[a.right] - real_prefix[a.left];
    if (nreal == 0) { // a leaf of ghosts only: nobody needs its list
        nrange[g] = 0;
        ncand[g]  = 0;
        return;
    }
    f64 a_rint = a.rint * Rkern;
    f64 e0x = a.lo0 - a_rint, e0y = a.lo1 - a_rint, e0z = a.lo2 - a_rint;
    f64 e1x = a.hi0 + a_rint, e1y = a.hi1 + a_rint, e1z = a.hi2 + a_rint;
    u32 stack[WALK_STACK];
    int sp      = 0;
    stack[sp++] = 0; // root (node 0; when I == 0 it is the only leaf)
    u32 nr = 0, cand = 0;
    u32 cur_s = 0xffffffffu, cur_e = 0xffffffffu;
    u32 *out = ranges + u64(g) * cap * 2;
    while (sp > 0) {
        u32 id     = stack[--sp];
        NodeRegs n = load_node(nodes + id);
        f64 r      = n.rint * Rkern;
        // cella_neigh_b(a, n ⊕ r) || cella_neigh_b(a ⊕ ra, n): fmax(x,y) <= fmin(u,v) ⟺ x<=v && y<=u when
        // x<=u and y<=v hold by construction (boxes are not inverted, r >= 0): same booleans, fewer FP64 ops
        bool hit = (a.lo0 <= n.hi0 + r && n.lo0 - r <= a.hi0 && a.lo1 <= n.hi1 + r && n.lo1 - r <= a.hi1
                    && a.lo2 <= n.hi2 + r && n.lo2 - r <= a.hi2)
                   || (e0x <= n.hi0 && n.lo0 <= e1x && e0y <= n.hi1 && n.lo1 <= e1y && e0z <= n.hi2 && n.lo2 <= e1z);
        if (!hit)
            continue;
        if (id >= I) { // leaf: ranks [left, right); DFS visits leaves in ascending order
            cand += n.right - n.left;
            if (n.left == cur_e) {
                cur_e = n.right;
            } else {
                if (cur_s != 0xffffffffu) {
                    if (nr < cap) {
                        out[2 * nr]     = cur_s;
                        out[2 * nr + 1] = cur_e;
                    }
                    nr++;
                }
                cur_s = n.left;
                cur_e = n.right;
            }
        } else {
            stack[sp++] = n.right;
            stack[sp++] = n.left;
        }
    }
    if (cur_s != 0xffffffffu) {
        if (nr < cap) {
            out[2 * nr]     = cur_s;
            out[2 * nr + 1] = cur_e;
        }
        nr++;
    }
    nrange[g] = nr < cap ? nr : cap; // clamped: an overflowing search is redone by the host
    ncand[g]  = cand;
    if (nr > cap)
        atomicMax(max_ranges, nr);
}

// ---------------------------------------------------------------------------------------------
// stage 2: accept pass (ballots) and ordered fill
// ---------------------------------------------------------------------------------------------
constexpr int S2_WARPS = 4;

struct WarpScratch {
    u32 start[RANGE_CAP_DEFAULT * 4];   // range starts (cap <= 256)
    u32 pre[RANGE_CAP_DEFAULT * 4 + 1]; // exclusive prefix of the range lengths
    Pack4 pa[32];                       // (x, y, z, lim_a) of the current batch of particles
};

/// loads the ranges of `leaf` into the warp scratch; returns the candidate count
__device__ __forceinline__ u32 load_ranges(
    WarpScratch &w, const u32 *__restrict__ ranges, u32 leaf, u32 cap, u32 nr, int lane) {
    u32 run = 0;
    for (u32 base = 0; base < nr; base += 32) {
        u32 k   = base + lane;
        u32 s = 0, len = 0;
        if (k < nr) {
            s   = ranges[(u64(leaf) * cap + k) * 2];
            len = ranges[(u64(leaf) * cap + k) * 2 + 1] - s;
        }
        u32 inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o)
                inc += t;
        }
        if (k < nr) {
            w.start[k] = s;
            w.pre[k]   = run + inc - len;
        }
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0)
        w.pre[nr] = run;
    __syncwarp();
    return run;
}
/// rank of candidate j (j < ncand): binary search of the range holding it
__device__ __forceinline__ u32 cand_rank(const WarpScratch &w, u32 nr, u32 j) {
    u32 lo = 0, hi = nr; // invariant: pre[lo] <= j < pre[hi]
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (w.pre[mid] <= j)
            lo = mid;
        else
            hi = mid;
    }
    return w.start[lo] + (j - w.pre[lo]);
}

constexpr u32 BALLOT_WORDS = 512; ///< per-warp shared ballot store: (particles of the batch) x (chunks)
constexpr u32 RANK_CACHE   = 512; ///< per-warp cache of candidate ranks (16 chunks)

struct WarpScratch2 {
    u32 ball[BALLOT_WORDS];
    u32 rankc[RANK_CACHE];
};

/// Stage 2, one launch: warp per leaf, lanes = candidates.  Pass 1 tests every (particle, candidate) pair
/// and keeps the ballots in shared memory; the warp then reserves the leaf's list space with one atomic
/// on a global cursor (lists are contiguous per leaf, leaves in completion order — the internal CSR does
/// not need a global order, the exported ObjectCache is rebuilt by id) and pass 2 replays the ballots.
/// Leaves whose ballots do not fit the shared store re-test in pass 2 instead.
__global__ void __launch_bounds__(S2_WARPS * 32) neigh_lists_kernel(
    const NodePack *__restrict__ nodes, u32 I, u32 L, const Pack4 *__restrict__ SA, const u8 *__restrict__ real_flag,
    const u32 *__restrict__ real_prefix, const u32 *__restrict__ ranges, u32 cap, const u32 *__restrict__ nrange,
    f64 Rker2, f64 h_tolerance, u64 list_cap, unsigned long long *__restrict__ cursor, u32 *__restrict__ cnt_s,
    u32 *__restrict__ off_s, u32 *__restrict__ list_s) {
    __shared__ WarpScratch ws[S2_WARPS];
    __shared__ WarpScratch2 ws2[S2_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 lt   = (1u << lane) - 1u;
    u32 leaf = blockIdx.x * S2_WARPS + warp;
    if (leaf >= L)
        return;
    u32 nr = nrange[leaf];
    if (nr == 0)
        return;
    WarpScratch &w   = ws[warp];
    WarpScratch2 &w2 = ws2[warp];
    const u32 p0 = reinterpret_cast<const u32 *>(&nodes[I + leaf].rint)[2];
    const u32 p1 = reinterpret_cast<const u32 *>(&nodes[I + leaf].rint)[3];
    const u32 ncand  = load_ranges(w, ranges, leaf, cap, nr, lane);
    const u32 nchunk = (ncand + 31) >> 5;
    const u32 slot0  = real_prefix[p0];
    const bool cache_ranks = nchunk * 32 <= RANK_CACHE;
    u32 a_base = 0;
    for (u32 rb = p0; rb < p1; rb += 32) {
        u32 r   = rb + lane;
        bool va = r < p1 && real_flag[r];
        u32 bal = __ballot_sync(0xffffffffu, va);
        u32 nb  = __popc(bal);
        if (nb == 0)
            continue;
        const bool keep_ballots = nb * nchunk <= BALLOT_WORDS;
        __syncwarp();
        if (va) {
            Pack4 q    = ld4(SA + r);
            f64 rint_a = q.d * h_tolerance;
            w.pa[__popc(bal & lt)] = Pack4{q.a, q.b, q.c, rint_a * rint_a * Rker2};
        }
        __syncwarp();
        // ---- pass 1: ballots + counts
        u32 mycount = 0;
        for (u32 c = 0; c < nchunk; c++) {
            u32 j   = c * 32 + lane;
            bool vb = j < ncand;
            f64 bx = 0, by = 0, bz = 0, lim_b = 0;
            if (vb) {
                u32 rank_b = cand_rank(w, nr, j);
                if (cache_ranks)
                    w2.rankc[j] = rank_b;
                Pack4 q = ld4(SA + rank_b);
                bx = q.a, by = q.b, bz = q.c;
                f64 rint_b = q.d * h_tolerance;
                lim_b      = rint_b * rint_b * Rker2;
            }
            u32 mymask = 0;
            for (u32 a = 0; a < nb; a++) {
                Pack4 pa = w.pa[a];
                f64 dx = pa.a - bx, dy = pa.b - by, dz = pa.c - bz;
                f64 rab2         = dx * dx + dy * dy + dz * dz;
                bool no_interact = rab2 > pa.d && rab2 > lim_b;
                u32 m            = __ballot_sync(0xffffffffu, vb && !no_interact);
                if (lane == int(a)) {
                    mymask = m;
                    mycount += __popc(m);
                }
            }
            if (keep_ballots && lane < int(nb))
                w2.ball[c * nb + lane] = mymask;
        }
        // ---- reserve the list space of this batch: exclusive prefix of the counts + one atomic
        u32 inc = lane < int(nb) ? mycount : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o)
                inc += t;
        }
        u32 total = __shfl_sync(0xffffffffu, inc, 31);
        unsigned long long base = 0;
        if (lane == 0)
            base = atomicAdd(cursor, (unsigned long long) total);
        base = __shfl_sync(0xffffffffu, base, 0);
        u32 myoff = u32(base) + inc - (lane < int(nb) ? mycount : 0u);
        if (lane < int(nb)) {
            cnt_s[slot0 + a_base + lane] = mycount;
            off_s[slot0 + a_base + lane] = myoff;
        }
        a_base += nb;
        if (base + total > list_cap)
            continue; // the list array is too small: the host reads the cursor and runs the search again
        __syncwarp();
        // ---- pass 2: ordered fill
        for (u32 c = 0; c < nchunk; c++) {
            u32 j      = c * 32 + lane;
            bool vb    = j < ncand;
            u32 rank_b = 0;
            if (vb)
                rank_b = cache_ranks ? w2.rankc[j] : cand_rank(w, nr, j);
            if (keep_ballots) {
                u32 mymask = lane < int(nb) ? w2.ball[c * nb + lane] : 0u;
                for (u32 a = 0; a < nb; a++) {
                    u32 m  = __shfl_sync(0xffffffffu, mymask, a);
                    u32 bo = __shfl_sync(0xffffffffu, myoff, a);
                    if ((m >> lane) & 1u)
                        list_s[bo + __popc(m & lt)] = rank_b;
                }
                myoff += __popc(mymask);
            } else { // ballots did not fit: test again
                f64 bx = 0, by = 0, bz = 0, lim_b = 0;
                if (vb) {
                    Pack4 q = ld4(SA + rank_b);
                    bx = q.a, by = q.b, bz = q.c;
                    f64 rint_b = q.d * h_tolerance;
                    lim_b      = rint_b * rint_b * Rker2;
                }
                for (u32 a = 0; a < nb; a++) {
                    Pack4 pa = w.pa[a];
                    f64 dx = pa.a - bx, dy = pa.b - by, dz = pa.c - bz;
                    f64 rab2         = dx * dx + dy * dy + dz * dz;
                    bool no_interact = rab2 > pa.d && rab2 > lim_b;
                    u32 m            = __ballot_sync(0xffffffffu, vb && !no_interact);
                    u32 bo           = __shfl_sync(0xffffffffu, myoff, a);
                    if ((m >> lane) & 1u)
                        list_s[bo + __popc(m & lt)] = rank_b;
                    if (lane == int(a))
                        myoff += __popc(m);
                }
            }
        }
    }
}

void search_build(
    cudaStream_t s, const TreeBuffers &tb, SearchBuffers &sb, const f64 *d_rint, f64 Rkern, f64 h_tolerance) {
    const u32 I = tb.I, L = tb.L;
    sb.nodes.ensure(size_t(I) + L, 1.1);
    pack_nodes_kernel<<<grid_for(size_t(I) + L, 256), 256, 0, s>>>(
        I, L, tb.aabb_min.p, tb.aabb_max.p, d_rint, tb.lchild.p, tb.rchild.p, tb.lflag.p, tb.rflag.p,
        tb.reduc_index_map.p, sb.nodes.p);
    SB_COUNT_LAUNCH();
    sb.nrange.ensure(L, 1.1);
    sb.ncand.ensure(L, 1.1);
    sb.cnt_s.ensure(sb.N, 1.1);
    sb.off_s.ensure(sb.N, 1.1);
    const f64 Rker2 = Rkern * Rkern;
    // list capacity: what the previous search of this patch needed (+ slack), ~96 entries per particle the
    // first time.  ONE host synchronisation per search: it reads the list cursor (= K) and the largest
    // range count; if either ran past its capacity the exact need is known and the search is repeated.
    if (sb.list_s.cap == 0)
        sb.list_s.ensure(size_t(sb.N) * 96 + 1024);
    u32 *d_maxr                  = reinterpret_cast<u32 *>(sb.scalars.p + 2);
    unsigned long long *d_cursor = reinterpret_cast<unsigned long long *>(sb.scalars.p + 3);
    for (int attempt = 0;; attempt++) {
        if (sb.range_cap > RANGE_CAP_DEFAULT * 4)
            throw std::runtime_error("neighbour search: more than 256 candidate rank ranges for one leaf");
        sb.ranges.ensure(size_t(L) * sb.range_cap * 2, 1.1);
        SB_CUDA_CHECK(cudaMemsetAsync(sb.scalars.p + 2, 0, 2 * sizeof(u64), s));
        leaf_ranges_kernel<<<grid_for(L, 128), 128, 0, s>>>(
            sb.nodes.p, I, L, sb.real_prefix.p, Rkern, sb.range_cap, sb.ranges.p, sb.nrange.p, sb.ncand.p, d_maxr);
        SB_COUNT_LAUNCH();
        neigh_lists_kernel<<<grid_for(L, S2_WARPS), S2_WARPS * 32, 0, s>>>(
            sb.nodes.p, I, L, sb.SA.p, sb.real_flag.p, sb.real_prefix.p, sb.ranges.p, sb.range_cap, sb.nrange.p, Rker2,
            h_tolerance, u64(sb.list_s.cap), d_cursor, sb.cnt_s.p, sb.off_s.p, sb.list_s.p);
        SB_COUNT_LAUNCH();
        SB_CUDA_CHECK(cudaMemcpyAsync(sb.h_scalars.p + 2, sb.scalars.p + 2, 2 * sizeof(u64), cudaMemcpyDeviceToHost, s));
        SB_CUDA_CHECK(cudaStreamSynchronize(s));
        u32 maxr = u32(sb.h_scalars.p[2] & 0xffffffffull);
        sb.K     = sb.h_scalars.p[3];
        bool redo = false;
        if (maxr > sb.range_cap) { // rare: a leaf sees more merged ranges than its slot holds
            while (sb.range_cap < maxr)
                sb.range_cap *= 2;
            redo = true;
        } else if (sb.K > 0xFFFFFFFFull) {
            throw std::overflow_error(
                "neighbour count overflows u32 (sum_neigh_cnt is u32 in the reference, TreeTraversal.hpp:378): "
                "use more / smaller patches");
        }
        if (sb.K > sb.list_s.cap && sb.K <= 0xFFFFFFFFull) {
            sb.list_s.ensure(sb.K, 1.1);
            redo = true;
        }
        if (!redo)
            break;
        if (attempt >= 3)
            throw std::runtime_error("neighbour search: capacity retry failed");
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
