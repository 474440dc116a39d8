// stream_kernels.cuh — launchers of the streaming kernels (stream_kernels.cu)
#pragma once
#include "sph.cuh"

namespace sb {
/// wrap_min / wrap_max (optional): the periodic box the drifted positions are wrapped into in the same pass
void leapfrog_predictor(
    cudaStream_t s, u32 n, f64 dt, f64 *xyz, f64 *vxyz, const f64 *axyz, f64 *uint_, const f64 *duint,
    const f64 *wrap_min = nullptr, const f64 *wrap_max = nullptr);
void leapfrog_predictor_pos(
    cudaStream_t s, u32 n, f64 dt, f64 *xyz, f64 *vxyz, const f64 *axyz, const f64 *wrap_min = nullptr,
    const f64 *wrap_max = nullptr);
void leapfrog_predictor_u(cudaStream_t s, u32 n, f64 dt, f64 *uint_, const f64 *duint);
void leapfrog_corrector(cudaStream_t s, u32 n, f64 hdt, f64 *vxyz, const f64 *axyz, const f64 *axyz_old, f64 *uint_,
                        const f64 *duint, const f64 *duint_old, u64 *red_max, f64 *red_sum, f64 *cons = nullptr);
void step_scalars(cudaStream_t s, const u64 *red, f64 *sc);
void periodic_wrap(cudaStream_t s, u32 n, f64 *xyz, const f64 bmin[3], const f64 bmax[3]);
void ext_force_point_mass(cudaStream_t s, u32 n, const f64 *xyz, f64 *axyz_ext, f64 central_mass, f64 G);
void flag_in_box(cudaStream_t s, u32 n, const f64 *xyz, const f64 lo[3], const f64 hi[3], u8 *flag);
void flag_sphere(cudaStream_t s, u32 n, const f64 *xyz, const f64 c[3], f64 rad, int mode, u8 *flag);
void patch_owner(cudaStream_t s, u32 n, const f64 *xyz, u32 npatch, const f64 *d_boxes, u32 self, u8 *stay_flag, u32 *owner);
void flag_equal(cudaStream_t s, u32 n, const u32 *v, u32 val, u8 *flag);
void scatter_ids(cudaStream_t s, u32 n, const u8 *flag, const u32 *pos, u32 *ids);
/// ghost selection of one sender patch against up to 64 cut boxes (d_boxes: nbox*6 doubles, [lo,hi))
void ghost_select_count(cudaStream_t s, u32 n, const f64 *xyz, u32 nbox, const f64 *d_boxes, u64 *mask, u32 *block_counts,
                        u32 *d_totals);
void ghost_select_scatter(cudaStream_t s, u32 n, const u64 *mask, u32 nbox, const u32 *block_offsets, const u64 *d_base,
                          u32 *ids_pool);
void gather_field(cudaStream_t s, u32 cnt, int nvar, const u32 *ids, const f64 *src, f64 *dst);
// ---- batched interface kernels -------------------------------------------------------------------------
// A patch has up to 26 interfaces and a rank many patches: a launch per interface and exchange costs more than
// the copies themselves (a few hundred ghosts each).  The jobs of up to BATCH_JOBS interfaces travel in the
// kernel parameters (a __grid_constant__ struct: no table in device memory, no copy) and one launch serves them
// all; a block finds its job by the first-block prefix.
constexpr int BATCH_JOBS = 40;
struct GhostXyzhJob { ///< A_dst[k] = (xyz[ids[k]] + offset, h[ids[k]]), ids == nullptr: identity
    const u32 *ids;
    const f64 *xyz, *h;
    Pack4 *dst;
    f64 ox, oy, oz;
    u32 count;
};
struct PackFieldsJob { ///< pack_fields of one interface / patch
    const u32 *ids, *dst_map;
    const f64 *h, *vxyz, *uint_, *omega, *axyz;
    Pack4 *A, *B, *C, *D;
    u32 count;
};
struct PackAlphaJob { ///< pack_alpha of one interface / patch
    const u32 *ids, *dst_map;
    const f64 *alpha, *omega;
    Pack4 *C;
    u32 count;
};
struct UnpackGhostJob { ///< unpack_ghost_fields of one received interface
    const Pack4 *sA, *sB, *sC, *sD;
    const u32 *dst_map;
    Pack4 *A, *B, *C, *D;
    u32 count;
};
template<class Job>
struct JobBatch {
    Job job[BATCH_JOBS];
    u32 first_block[BATCH_JOBS + 1];
    int n;
};
/// collects jobs and launches one kernel per BATCH_JOBS of them (and at flush / destruction)
template<class Job>
struct Batcher {
    cudaStream_t s;
    JobBatch<Job> b;
    u32 blocks = 0;
    explicit Batcher(cudaStream_t st) : s(st) { b.n = 0; }
    void add(const Job &j) {
        if (!j.count)
            return;
        b.job[b.n]         = j;
        b.first_block[b.n] = blocks;
        blocks += (j.count + 255) / 256;
        if (++b.n == BATCH_JOBS)
            flush();
    }
    void flush();
    ~Batcher() { flush(); }
};

/// every field of a patch data layout in ONE launch: row k of the destination (from row dst_off on) = row
/// ids[k] (ids == nullptr: src_off + k) of the source, field by field
struct RowTable {
    const f64 *src[12];
    f64 *dst[12];
    int nvar[12];
    int nf;
};
void rows_gather(cudaStream_t s, u32 cnt, const u32 *ids, const RowTable &t, u32 src_off, u32 dst_off);
/// pos = exclusive scan of flag: set[pos[i]] = i where flag[i], cleared[i - pos[i]] = i elsewhere (both ascending)
void split_ids(cudaStream_t s, u32 n, const u8 *flag, const u32 *pos, u32 *set_ids, u32 *cleared_ids);
/// out = the ids[j], in order, whose key[ids[j]] == val (one block: meant for the few objects that change patch)
void select_equal(cudaStream_t s, u32 n, const u32 *ids, const u32 *key, u32 val, u32 *out);
void pack_xyzh(cudaStream_t s, u32 n, const f64 *xyz, const f64 *h, Pack4 *A);
void ghost_xyzh(cudaStream_t s, u32 cnt, const u32 *ids, const f64 *xyz, const f64 *h, const f64 off[3], Pack4 *A_dst);
/// dst_map / src_map (optional) redirect record k to slot map[k] (Morton-sorted storage)
void pack_fields(cudaStream_t s, u32 cnt, const u32 *ids, const f64 *h, const f64 *vxyz, const f64 *uint_,
                 const f64 *omega, const f64 *axyz, Pack4 *A, Pack4 *B, Pack4 *C, Pack4 *D, const u32 *dst_map = nullptr);
void unpack_ghost_fields(cudaStream_t s, u32 cnt, const Pack4 *sA, const Pack4 *sB, const Pack4 *sC, const Pack4 *sD,
                         Pack4 *A, Pack4 *B, Pack4 *C, Pack4 *D, const u32 *dst_map = nullptr);
void pack_alpha(cudaStream_t s, u32 cnt, const u32 *ids, const f64 *alpha, Pack4 *C, const u32 *dst_map = nullptr,
                const f64 *omega = nullptr);
void unpack_cs(cudaStream_t s, u32 n, const Pack4 *C, f64 *cs, const u32 *src_map = nullptr);
void unpack_comp(cudaStream_t s, u32 n, const Pack4 *P, int first, int nc, f64 *out, const u32 *src_map = nullptr);
void max_reduce(cudaStream_t s, u32 n, const f64 *v, u64 *red);
/// *out = number of objects whose successor by id lies farther away than 8 h (locality of the id order)
void count_far_successors(cudaStream_t s, u32 n, const f64 *xyz, const f64 *h, u64 *out);
} // namespace sb
