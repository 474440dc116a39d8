// neigh2.cuh — the B200 neighbour search: Morton-sorted particle storage, packed tree nodes, one
// warp-cooperative tree walk per group of 8 leaves (candidate leaves + member masks), a warp-cooperative
// accept pass whose lanes are the candidates (ballot masks), and an ordered fill.  Output: a CSR over the REAL particles in sorted
// (Morton rank) order whose entries are RANKS; `export_object_cache` converts it to the reference's
// ObjectCache layout (by particle id, ids) — bit-identical to NeighbourCache.cpp:223-604.
#pragma once
#include "sph.cuh"
#include "tree.cuh"
#include <functional>

namespace sb {

/// one tree node in 64 bytes (two 32-byte sectors): box, interaction radius, children / leaf range
struct alignas(64) NodePack {
    f64 lo[3], hi[3];
    f64 rint;       ///< max(h) * htol of the node's objects
    u32 left, right; ///< internal: child node ids (leaves are offset by I); leaf: rank range [left, right)
};

constexpr int GL = 8; ///< leaves per walk group (the member mask lives in the top 8 bits of an entry)

struct SearchBuffers {
    u32 N = 0, M = 0, L = 0, I = 0;
    u64 K        = 0; ///< total neighbour count
    u64 pair_tests = 0; ///< (particle, candidate) accept tests of the last search
    u64 entries_last = 0; ///< candidate entries the last search needed
    u32 frontier_cap = 320; ///< walk frontier entries per group in shared memory (doubled on demand)
    // how the last search went (shamb200_neigh_cache_stats): attempts (1 = every capacity was large enough),
    // groups that were walked with a frontier in global memory
    u32 attempts_last = 0, over_groups_last = 0;
    DevBuf<NodePack> nodes;  // [I+L]
    DevBuf<Pack4> SA;        // [M] (x,y,z,h) in sorted order
    DevBuf<u32> inv_map;     // [M] rank of merged index i
    DevBuf<u8> real_flag;    // [M+1]
    DevBuf<u32> real_prefix; // [M+1] exclusive scan of real_flag over ranks (slot of a real rank)
    DevBuf<u32> slot_rank;   // [N] rank of slot k
    DevBuf<uint2> gcand;     // (first rank, member mask << 24 | length) per candidate leaf, group after group
    DevBuf<u64> gc_off;      // [G] first entry of group g
    DevBuf<u32> gcount;      // [G] number of entries of group g
    DevBuf<u32> over_list;   // groups whose walk did not fit the shared-memory frontier
    DevBuf<uint2> big_scratch; // their frontiers in global memory
    DevBuf<u32> top_front, top_count; // start frontiers of the group walks, one per 64 groups
    DevBuf<u32> cnt_s, off_s; // [N]
    DevBuf<u32> list_s;       // [K] ranks, ascending inside each list; lists of one leaf contiguous
    DevBuf<u32> scan_tmp;
    DevBuf<u64> scalars;
    PinnedBuf<u64> h_scalars;
    // exported ObjectCache (by id), built on demand
    bool exported = false;
    DevBuf<u32> x_cnt, x_scanned, x_list;
};

/// rank-CSR view handed to the SPH loops.  A launch handles `count` real particles starting at `first`:
/// slots [first, first + count) in Morton order (the default: neighbouring lanes share their neighbours), or —
/// by_id — the objects with ids [first, first + count), whatever their slots: the host-resident step runs its
/// loops over id ranges so that the outputs of a finished range (contiguous in the by-id fields) can travel to
/// the host while the next range is computed (Model::evolve_once_host)
struct RankCsr {
    const u32 *cnt, *off, *list; ///< by slot; entries are ranks
    const u32 *slot_rank;        ///< [N]
    const u32 *index_map;        ///< rank -> merged id
    u32 N;
    u32 first, count;
    bool by_id;
    const u32 *inv_map;     ///< merged id -> rank
    const u32 *real_prefix; ///< rank -> slot (real objects)
    RankCsr ids(u32 id0, u32 n) const {
        RankCsr c = *this;
        c.first = id0, c.count = n, c.by_id = true;
        return c;
    }
};
inline RankCsr rank_csr_of(const SearchBuffers &sb, const TreeBuffers &tb) {
    return RankCsr{sb.cnt_s.p, sb.off_s.p, sb.list_s.p, sb.slot_rank.p, tb.index_map.p, sb.N,
                   0u,         sb.N,       false,       sb.inv_map.p,   sb.real_prefix.p};
}
#ifdef __CUDACC__
/// work item k (< c.count) of a launch -> slot kk, rank r, id
__device__ __forceinline__ void csr_item(const RankCsr &c, u32 k, u32 &kk, u32 &r, u32 &id) {
    if (c.by_id) {
        id = c.first + k;
        r  = c.inv_map[id];
        kk = c.real_prefix[r];
    } else {
        kk = c.first + k;
        r  = c.slot_rank[kk];
        id = c.index_map[r];
    }
}
#endif

/// SA[r] = A[index_map[r]], inv_map, real flags / slots.  A: merged (x,y,z,h) by id, N real objects first.
void search_prepare_sorted(cudaStream_t s, const TreeBuffers &tb, SearchBuffers &sb, const Pack4 *A, u32 N);
/// same from separate position / h arrays (stage-level C ABI)
void search_prepare_sorted_strided(
    cudaStream_t s, const TreeBuffers &tb, SearchBuffers &sb, const f64 *xyz, size_t stride, const f64 *h,
    size_t hstride, u32 N);
/// the search proper (two-stage criterion of the reference).  Synchronises once (list sizing).
/// `mark(name)`, when given, is called on the stream before the walk ("neigh_walk") and before the list
/// kernel ("neigh_lists"): stage timing of the caller.
void search_build(
    cudaStream_t s, const TreeBuffers &tb, SearchBuffers &sb, const f64 *d_rint, f64 Rkern, f64 h_tolerance,
    const std::function<void(const char *)> &mark = {});
/// the same around ONE stream synchronisation of the caller (several patches: enqueue all, synchronise once,
/// finish all).  `finish` repeats the search of a patch whose capacities were exceeded (synchronising itself).
void search_enqueue(
    cudaStream_t s, const TreeBuffers &tb, SearchBuffers &sb, const f64 *d_rint, f64 Rkern, f64 h_tolerance);
void search_finish(cudaStream_t s, const TreeBuffers &tb, SearchBuffers &sb, f64 Rkern, f64 h_tolerance);
/// ObjectCache layout of the reference (cnt / scanned by id, ids).  Synchronises.
void export_object_cache(cudaStream_t s, const TreeBuffers &tb, SearchBuffers &sb);

} // namespace sb
