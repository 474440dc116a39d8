// solver.cu — Solver::evolve_once on the GPU: the host-side driver of the B200 SPH step.
//
// Follows shammodels/sph/src/Solver.cpp:1942-3272 (evolve_once), :1060-1304 (sph_prestep),
// :1394-1633 (communicate_merge_ghosts_fields), shammodels/sph/src/BasicSPHGhosts.cpp:261-579 and
// shammodels/sph/include/shammodels/sph/{BasicSPHGhosts,SPHUtilities}.hpp (ghost zones) — paths
// relative to /root/reference/src.  Every device operation is a hand-written kernel of this
// library; there is no CPU fallback for any stage.
#include "solver.cuh"
#include <cstdlib>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>

namespace sb {


// ---------------------------------------------------------------------------------------------
// PatchFields
// ---------------------------------------------------------------------------------------------
std::vector<PatchFields::Ref> PatchFields::all() {
    return {{"xyz", &xyz, 3},         {"vxyz", &vxyz, 3},   {"axyz", &axyz, 3},
            {"axyz_ext", &axyz_ext, 3}, {"hpart", &hpart, 1}, {"uint", &uint_, 1},
            {"duint", &duint, 1},     {"alpha_AV", &alpha_AV, 1}, {"divv", &divv, 1},
            {"dtdivv", &dtdivv, 1},   {"curlv", &curlv, 3}, {"soundspeed", &soundspeed, 1}};
}
void PatchFields::reserve(u32 cap, cudaStream_t s) {
    for (auto &r : all()) {
        const size_t old_cap = r.buf->cap;
        r.buf->ensure_keep(size_t(cap) * r.nvar, size_t(n) * r.nvar, s, 1.25);
        if (r.buf->cap != old_cap) // a fresh block: fields start at 0 like the reference's PatchDataField
            SB_CUDA_CHECK(cudaMemsetAsync(
                r.buf->p + size_t(n) * r.nvar, 0, (r.buf->cap - size_t(n) * r.nvar) * sizeof(f64), s));
    }
}

// ---------------------------------------------------------------------------------------------
// StageTimer
// ---------------------------------------------------------------------------------------------
void StageTimer::begin_step() { marks.clear(); }
void StageTimer::mark(cudaStream_t s, const char *name) {
    int idx = int(marks.size());
    if (idx >= int(pool.size())) {
        cudaEvent_t e;
        SB_CUDA_CHECK(cudaEventCreate(&e));
        pool.push_back(e);
    }
    SB_CUDA_CHECK(cudaEventRecord(pool[idx], s));
    marks.emplace_back(name, idx);
}
void StageTimer::end_step(cudaStream_t s) {
    mark(s, "end");
    SB_CUDA_CHECK(cudaEventSynchronize(pool[marks.back().second]));
    acc.clear();
    order.clear();
    for (size_t k = 0; k + 1 < marks.size(); k++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, pool[marks[k].second], pool[marks[k + 1].second]);
        if (!acc.count(marks[k].first))
            order.push_back(marks[k].first);
        acc[marks[k].first] += ms;
    }
    names_joined.clear();
    values.clear();
    for (auto &n : order) {
        if (!names_joined.empty())
            names_joined += ";";
        names_joined += n;
        values.push_back(acc[n]);
    }
}
StageTimer::~StageTimer() {
    for (auto e : pool)
        cudaEventDestroy(e);
}

// ---------------------------------------------------------------------------------------------
// small device helpers local to the solver
// ---------------------------------------------------------------------------------------------
/// red slots: 0 eps max, 1 eps min, 2 eps_v² max, 3 Σv² (f64 bits), 4 cfl min, 8.. per-patch h max
__global__ void reset_red_kernel(u64 *red, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        u64 v = 0; // max accumulators (ordered encoding: 0 is below every double) and the f64 sum
        if (i == 1 || i == 4)
            v = 0xFFFFFFFFFFFFFFFFull; // min accumulators
        red[i] = v;
    }
}
/// 8 step scalars + one slot per patch (max h)
static int red_slots(const Model &m) { return 8 + int(std::max<size_t>(256, m.patches.size())); }

void Model::reset_red() {
    const int n = red_slots(*this);
    red.ensure(n);
    h_red.ensure(n);
    reset_red_kernel<<<1, 512, 0, s()>>>(red.p, n);
    SB_COUNT_LAUNCH();
}
void Model::read_red(int n) {
    d2h_small(s(), h_red.p, red.p, size_t(n) * sizeof(u64));
    SB_CUDA_CHECK(cudaStreamSynchronize(s()));
}
static f64 bits_to_f64(u64 b) {
    f64 d;
    std::memcpy(&d, &b, 8);
    return d;
}

/// counts[b] = number of particles inside box b ([lo,hi) per axis); boxes: nb*6 doubles
__global__ void __launch_bounds__(256) count_in_boxes_kernel(
    u32 n, const f64 *__restrict__ xyz, u32 nb, const f64 *__restrict__ boxes, u32 *__restrict__ counts) {
    extern __shared__ u32 sc[];
    for (u32 j = threadIdx.x; j < nb; j += blockDim.x)
        sc[j] = 0;
    __syncthreads();
    for (u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += u64(gridDim.x) * blockDim.x) {
        f64 x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        for (u32 b = 0; b < nb; b++) {
            const f64 *q = boxes + 6 * b;
            if (q[0] <= x && x < q[3] && q[1] <= y && y < q[4] && q[2] <= z && z < q[5])
                atomicAdd(&sc[b], 1u);
        }
    }
    __syncthreads();
    for (u32 j = threadIdx.x; j < nb; j += blockDim.x)
        if (sc[j])
            atomicAdd(&counts[j], sc[j]);
}

// ---------------------------------------------------------------------------------------------
// setup
// ---------------------------------------------------------------------------------------------
void Model::set_box(const f64 bmin[3], const f64 bmax[3], u32 nx, u32 ny, u32 nz) {
    for (int d = 0; d < 3; d++) {
        box_min[d] = bmin[d];
        box_max[d] = bmax[d];
    }
    std::vector<PatchBox> grid = plan_patch_grid(bmin, bmax, nx, ny, nz, world);
    patches.clear();
    patches.resize(grid.size());
    for (size_t k = 0; k < grid.size(); k++)
        static_cast<PatchBox &>(patches[k]) = grid[k];
    u32 np = u32(grid.size());
    next_patch_id = np;
    std::vector<f64> hb(size_t(np) * 6);
    for (u32 k = 0; k < np; k++)
        for (int d = 0; d < 3; d++) {
            hb[6 * k + d]     = patches[k].lo[d];
            hb[6 * k + 3 + d] = patches[k].hi[d];
        }
    d_boxes.ensure(hb.size());
    h2d_small(s(), d_boxes.p, hb.data(), hb.size() * sizeof(f64));
    SB_CUDA_CHECK(cudaStreamSynchronize(s()));
}

void Model::push_particles(u64 n, const f64 *xyz, const f64 *vxyz, const f64 *h, const f64 *u, const f64 *alpha) {
    if (patches.empty())
        throw std::runtime_error("the box size is not set, please resize the box to the domain size");
    // host-side binning (setup path, not the hot path)
    size_t np = patches.size();
    std::vector<std::vector<u64>> bins(np);
    for (u64 i = 0; i < n; i++) {
        int own = -1;
        for (size_t k = 0; k < np; k++) {
            const PatchD &p = patches[k];
            if (p.lo[0] <= xyz[3 * i] && xyz[3 * i] < p.hi[0] && p.lo[1] <= xyz[3 * i + 1]
                && xyz[3 * i + 1] < p.hi[1] && p.lo[2] <= xyz[3 * i + 2] && xyz[3 * i + 2] < p.hi[2]) {
                own = int(k);
                break;
            }
        }
        if (own < 0) {
            if (np == 1)
                own = 0;
            else
                throw std::runtime_error("particle outside of the simulation box");
        }
        bins[own].push_back(i);
    }
    for (size_t k = 0; k < np; k++) {
        PatchD &p = patches[k];
        if (!is_local(p) || bins[k].empty())
            continue;
        u32 add  = u32(bins[k].size());
        u32 newn = p.f.n + add;
        p.f.reserve(newn, s());
        auto up = [&](DevBuf<f64> &buf, int nv, const f64 *src) {
            host_tmp.assign(size_t(add) * nv, 0.);
            if (src)
                for (u32 q = 0; q < add; q++)
                    for (int c = 0; c < nv; c++)
                        host_tmp[size_t(q) * nv + c] = src[bins[k][q] * nv + c];
            SB_CUDA_CHECK(cudaMemcpyAsync(
                buf.p + size_t(p.f.n) * nv, host_tmp.data(), host_tmp.size() * sizeof(f64),
                cudaMemcpyHostToDevice, s()));
            SB_CUDA_CHECK(cudaStreamSynchronize(s()));
        };
        for (auto &r : p.f.all()) {
            const f64 *src = nullptr;
            std::string nm(r.name);
            if (nm == "xyz")
                src = xyz;
            else if (nm == "vxyz")
                src = vxyz;
            else if (nm == "hpart")
                src = h;
            else if (nm == "uint")
                src = u;
            else if (nm == "alpha_AV")
                src = alpha;
            up(*r.buf, r.nvar, src);
        }
        p.f.n = newn;
    }
}

void Model::set_field(u32 ip, const std::string &name, const f64 *in, u64 count) {
    PatchD &p = patches.at(ip);
    if (!is_local(p))
        throw std::invalid_argument("patch is not local");
    for (auto &r : p.f.all())
        if (name == r.name) {
            if (count != u64(p.f.n) * r.nvar)
                throw std::invalid_argument("field size mismatch for " + name);
            SB_CUDA_CHECK(cudaMemcpyAsync(r.buf->p, in, count * sizeof(f64), cudaMemcpyHostToDevice, s()));
            SB_CUDA_CHECK(cudaStreamSynchronize(s()));
            return;
        }
    throw std::invalid_argument("unknown field " + name);
}

// ---------------------------------------------------------------------------------------------
// field access (ctx.collect_data equivalent + step internals for the parity tests)
// ---------------------------------------------------------------------------------------------
template<class T>
static int64_t copy_dev(cudaStream_t s, const T *d, size_t n, void *out, int64_t cap) {
    int64_t nb = int64_t(n * sizeof(T));
    if (out && cap >= nb && nb > 0) {
        SB_CUDA_CHECK(cudaMemcpyAsync(out, d, size_t(nb), cudaMemcpyDeviceToHost, s));
        SB_CUDA_CHECK(cudaStreamSynchronize(s));
    }
    return nb;
}

int64_t Model::get(u32 ip, const std::string &name, void *out, int64_t cap) {
    if (ip >= patches.size())
        return -1;
    PatchD &p = patches[ip];
    if (!is_local(p))
        return 0;
    for (auto &r : p.f.all())
        if (name == r.name)
            return copy_dev(s(), r.buf->p, size_t(p.f.n) * r.nvar, out, cap);
    PatchStep &st = p.st;
    auto packc    = [&](const Pack4 *P, u32 cnt, int first, int nc, const u32 *map) -> int64_t {
        int64_t nb = int64_t(size_t(cnt) * nc * sizeof(f64));
        if (out && cap >= nb && nb > 0) {
            field_tmp.ensure(size_t(cnt) * nc);
            unpack_comp(s(), cnt, P, first, nc, field_tmp.p, map);
            return copy_dev(s(), field_tmp.p, size_t(cnt) * nc, out, cap);
        }
        return nb;
    };
    const u32 *inv = st.srch.inv_map.p;
    if (name == "step.mxyz") return packc(st.A.p, st.m, 0, 3, nullptr);
    if (name == "step.mh") return cfg.keep_step_data ? copy_dev(s(), st.mh_snapshot.p, st.m, out, cap) : -1;
    if (name == "step.g_h") return packc(st.srch.SA.p, st.m, 3, 1, inv);
    if (name == "step.g_v") return packc(st.SB.p, st.m, 0, 3, inv);
    if (name == "step.g_u") return packc(st.SB.p, st.m, 3, 1, inv);
    if (name == "step.pressure") return packc(st.SC.p, st.m, 0, 1, inv);
    if (name == "step.g_omega") return packc(st.SC.p, st.m, 1, 1, inv);
    if (name == "step.soundspeed") return packc(st.SC.p, st.m, 2, 1, inv);
    if (name == "step.g_alpha") return packc(st.SC.p, st.m, 3, 1, inv);
    if (name == "step.g_a") return packc(st.SD.p, st.m, 0, 3, inv);
    if (name == "step.rint") return copy_dev(s(), st.rint.p, size_t(st.tree.I) + st.tree.L, out, cap);
    if (name == "step.omega") return copy_dev(s(), st.omega.p, st.n, out, cap);
    if (name == "step.alpha_updated") return copy_dev(s(), st.alpha_updated.p, st.n, out, cap);
    if (name == "step.vsig") return copy_dev(s(), st.vsig.p, st.n, out, cap);
    if (name == "step.cfl_dt") return copy_dev(s(), st.cfl_dt.p, st.n, out, cap);
    const TreeBuffers &t = st.tree;
    if (name == "tree.sorted_morton") return copy_dev(s(), t.morton.p, t.P2, out, cap);
    if (name == "tree.sort_index_map") return copy_dev(s(), t.index_map.p, t.P2, out, cap);
    if (name == "tree.reduc_index_map") return copy_dev(s(), t.reduc_index_map.p, size_t(t.L) + 2, out, cap);
    if (name == "tree.reduced_morton") return copy_dev(s(), t.reduced_morton.p, t.L, out, cap);
    if (name == "tree.lchild_id") return copy_dev(s(), t.lchild.p, t.I, out, cap);
    if (name == "tree.rchild_id") return copy_dev(s(), t.rchild.p, t.I, out, cap);
    if (name == "tree.lchild_flag") return copy_dev(s(), t.lflag.p, t.I, out, cap);
    if (name == "tree.rchild_flag") return copy_dev(s(), t.rflag.p, t.I, out, cap);
    if (name == "tree.endrange") return copy_dev(s(), t.endrange.p, t.I, out, cap);
    if (name == "tree.aabb_min") return copy_dev(s(), t.aabb_min.p, (size_t(t.I) + t.L) * 3, out, cap);
    if (name == "tree.aabb_max") return copy_dev(s(), t.aabb_max.p, (size_t(t.I) + t.L) * 3, out, cap);
    if (name.rfind("cache.", 0) == 0) { // the reference's ObjectCache layout, converted on demand
        export_object_cache(s(), st.tree, st.srch);
        if (name == "cache.cnt_neigh") return copy_dev(s(), st.srch.x_cnt.p, st.srch.N, out, cap);
        if (name == "cache.scanned_cnt") return copy_dev(s(), st.srch.x_scanned.p, st.srch.N, out, cap);
        if (name == "cache.index_neigh_map") return copy_dev(s(), st.srch.x_list.p, st.srch.K, out, cap);
    }
    return -1;
}

// ---------------------------------------------------------------------------------------------
// particle removal / migration (order preserving)
// ---------------------------------------------------------------------------------------------
/// doubles per object of the main layout (Σ nvar of PatchFields::all())
constexpr int kRowDoubles = 22;
/// one f64 block holding every field of `cnt` rows, field after field (field r starts at cnt * Σ nvar before it)
static RowTable fields_to_block(PatchFields &f, f64 *block, u32 cnt) {
    RowTable t{};
    size_t o = 0;
    int k    = 0;
    for (auto &r : f.all()) {
        t.src[k] = r.buf->p, t.dst[k] = block + o, t.nvar[k] = r.nvar;
        o += size_t(cnt) * r.nvar;
        k++;
    }
    t.nf = k;
    return t;
}
static RowTable block_to_fields(const f64 *block, u32 cnt, PatchFields &f) {
    RowTable t{};
    size_t o = 0;
    int k    = 0;
    for (auto &r : f.all()) {
        t.src[k] = block + o, t.dst[k] = r.buf->p, t.nvar[k] = r.nvar;
        o += size_t(cnt) * r.nvar;
        k++;
    }
    t.nf = k;
    return t;
}

/// keep the particles whose flag is set, in order (PatchDataLayer::keep_ids semantics); flg: the patch's flags
/// (default: the shared scratch `flag`)
void Model::keep_flagged(PatchD &p, u32 *out_kept, const u8 *flg) {
    u32 n = p.f.n;
    if (!flg)
        flg = flag.p;
    pos.ensure(n);
    exclusive_scan<u8>(s(), flg, pos.p, n, scan_tmp, red.p + 5);
    d2h_small(s(), h_red.p + 5, red.p + 5, sizeof(u64));
    SB_CUDA_CHECK(cudaStreamSynchronize(s()));
    u32 kept = u32(h_red.p[5]);
    if (out_kept)
        *out_kept = kept;
    if (kept == n)
        return;
    owner_tmp.ensure(n);
    scatter_ids(s(), n, flg, pos.p, owner_tmp.p);
    // all fields in two launches: kept rows -> one block -> back to the front of the fields
    field_tmp.ensure(size_t(kept) * kRowDoubles + 1);
    rows_gather(s(), kept, owner_tmp.p, fields_to_block(p.f, field_tmp.p, kept), 0, 0);
    rows_gather(s(), kept, nullptr, block_to_fields(field_tmp.p, kept, p.f), 0, 0);
    p.f.n = kept;
}

/// number of cleared flags of a patch, added to *out
__global__ void __launch_bounds__(256) count_cleared_kernel(u32 n, const u8 *__restrict__ flag, u32 *__restrict__ out) {
    u32 c = 0;
    for (u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += u64(gridDim.x) * blockDim.x)
        c += flag[i] ? 0u : 1u;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c)
        atomicAdd(out, c);
}

/// sphere removals of every local patch (accretion / kill spheres) with ONE synchronisation: the flags of all
/// patches are set and counted first; only a patch that really loses objects is compacted (keep_flagged)
void Model::remove_in_sphere(const f64 center[3], f64 radius, int mode) {
    std::vector<size_t> loc;
    std::vector<u64> off;
    u64 total = 0;
    for (size_t k = 0; k < patches.size(); k++)
        if (is_local(patches[k]) && patches[k].f.n) {
            loc.push_back(k);
            off.push_back(total);
            total += (u64(patches[k].f.n) + 15) / 16 * 16;
        }
    if (loc.empty())
        return;
    flag.ensure(total);
    box_counts.ensure(loc.size() + 1);
    SB_CUDA_CHECK(cudaMemsetAsync(box_counts.p, 0, (loc.size() + 1) * sizeof(u32), s()));
    for (size_t q = 0; q < loc.size(); q++) {
        PatchD &p = patches[loc[q]];
        flag_sphere(s(), p.f.n, p.f.xyz.p, center, radius, mode, flag.p + off[q]);
        const unsigned nb = (unsigned) std::min<u64>(u64(kNumSM) * 4, (u64(p.f.n) + 255) / 256);
        count_cleared_kernel<<<nb, 256, 0, s()>>>(p.f.n, flag.p + off[q], box_counts.p + q);
        SB_COUNT_LAUNCH();
    }
    h_counts.ensure(loc.size() + 1);
    d2h_small(s(), h_counts.p, box_counts.p, loc.size() * sizeof(u32));
    SB_CUDA_CHECK(cudaStreamSynchronize(s()));
    std::vector<u32> gone(h_counts.p, h_counts.p + loc.size());
    for (size_t q = 0; q < loc.size(); q++)
        if (gone[q])
            keep_flagged(patches[loc[q]], nullptr, flag.p + off[q]);
}

/// modules::ParticleReordering::reorder_particles (shammodels/sph/src/modules/ParticleReordering.cpp:22-51):
/// per patch, the Morton order of the positions over the PATCH box (RadixTreeMortonBuilder.cpp:68-107) and
/// PatchDataLayer::index_remap — new[i] = old[index_map[i]] for every field of the main layout
void Model::reorder_particles() {
    SB_CUDA_CHECK(cudaSetDevice(ctx->device));
    for (auto &p : patches) {
        if (!is_local(p) || !p.f.n)
            continue;
        const u32 n = p.f.n;
        morton_sort_permutation(s(), reorder_tree, p.f.xyz.p, 3, n, p.lo, p.hi, cfg.sort_mode);
        for (auto &r : p.f.all()) {
            field_tmp.ensure(size_t(n) * r.nvar + 1);
            gather_field(s(), n, r.nvar, reorder_tree.index_map.p, r.buf->p, field_tmp.p);
            SB_CUDA_CHECK(cudaMemcpyAsync(
                r.buf->p, field_tmp.p, size_t(n) * r.nvar * sizeof(f64), cudaMemcpyDeviceToDevice, s()));
        }
    }
}

/// ExternalForces::point_mass_accrete_particles (ExternalForces.cpp:593-700)
void Model::point_mass_accrete_particles() {
    if (!cfg.has_point_mass)
        return;
    const f64 c[3] = {0, 0, 0};
    remove_in_sphere(c, cfg.pm_racc, 0);
}
/// "part killing step" (Solver.cpp:526-578)
void Model::kill_particles() {
    for (int k = 0; k < cfg.n_kill_spheres; k++)
        remove_in_sphere(cfg.kill_center[k], cfg.kill_radius[k], 1);
}
/// ExternalForces::compute_ext_forces_indep_v (ExternalForces.cpp:49-323)
void Model::compute_ext_forces_indep_v() {
    for (auto &p : patches) {
        if (!is_local(p) || !p.f.n)
            continue;
        SB_CUDA_CHECK(cudaMemsetAsync(p.f.axyz_ext.p, 0, size_t(p.f.n) * 3 * sizeof(f64), s()));
        if (cfg.has_point_mass)
            ext_force_point_mass(s(), p.f.n, p.f.xyz.p, p.f.axyz_ext.p, cfg.pm_mass, cfg.constant_G);
    }
}

/// Solver::apply_position_boundary (Solver.cpp:938-981)
void Model::apply_position_boundary() {
    if (cfg.bc == SHAMB200_BC_PERIODIC && !wrapped_in_drift)
        for (auto &p : patches)
            if (is_local(p) && p.f.n)
                periodic_wrap(s(), p.f.n, p.f.xyz.p, box_min, box_max);
    wrapped_in_drift = false;
    reattribute_patch_objects();
}

/// histogram of the owner patch ids of one patch's particles (slot np = "no owner")
__global__ void __launch_bounds__(256) owner_hist_kernel(u32 n, const u32 *__restrict__ owner, u32 np, u32 *__restrict__ counts) {
    extern __shared__ u32 sc[];
    for (u32 j = threadIdx.x; j <= np; j += blockDim.x)
        sc[j] = 0;
    __syncthreads();
    for (u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += u64(gridDim.x) * blockDim.x) {
        u32 o = owner[i];
        atomicAdd(&sc[o < np ? o : np], 1u);
    }
    __syncthreads();
    for (u32 j = threadIdx.x; j <= np; j += blockDim.x)
        if (sc[j])
            atomicAdd(&counts[j], sc[j]);
}

/// ReattributeDataUtility::reatribute_patch_objects (ReattributeDataUtility.hpp:40-230): particles that
/// left their patch are extracted (order preserving) and appended to their new owner, senders in
/// ascending patch id (the multimap order of the reference's part_exchange).  Across ranks the rows
/// travel as one NCCL send/recv per field.
void Model::reattribute_patch_objects() {
    const size_t np = patches.size();
    if (np == 1)
        return; // single patch: periodic → everything is inside after the wrap; free → no constraint
    // 1. owners + per-destination counts of every local patch: all histograms land in one device array and come
    //    back with ONE copy and ONE synchronisation (a read-back per patch costs 0.4 ms of latency each)
    std::vector<DevBuf<u32>> owners(np);
    std::vector<u64> cm(np * np, 0); // cm[src*np + dst]
    std::vector<size_t> loc;
    for (size_t k = 0; k < np; k++)
        if (is_local(patches[k]) && patches[k].f.n)
            loc.push_back(k);
    box_counts.ensure(loc.size() * (np + 1) + 1);
    SB_CUDA_CHECK(cudaMemsetAsync(box_counts.p, 0, (loc.size() * (np + 1) + 1) * sizeof(u32), s()));
    const size_t hist_smem = (np + 1) * sizeof(u32);
    if (hist_smem > 48 * 1024)
        throw std::runtime_error("reattribute_patch_objects: too many patches for the owner histogram");
    for (size_t q = 0; q < loc.size(); q++) {
        const size_t k = loc[q];
        PatchD &p      = patches[k];
        flag.ensure(p.f.n);
        owners[k].ensure(p.f.n);
        patch_owner(s(), p.f.n, p.f.xyz.p, u32(np), d_boxes.p, u32(k), flag.p, owners[k].p);
        unsigned nb = (unsigned) std::min<u64>(u64(kNumSM) * 8, (u64(p.f.n) + 255) / 256);
        owner_hist_kernel<<<nb, 256, hist_smem, s()>>>(p.f.n, owners[k].p, u32(np), box_counts.p + q * (np + 1));
        SB_COUNT_LAUNCH();
    }
    {
        const size_t nhc = loc.size() * (np + 1) + 1;
        h_counts.ensure(nhc);
        d2h_small(s(), h_counts.p, box_counts.p, nhc * sizeof(u32));
        SB_CUDA_CHECK(cudaStreamSynchronize(s()));
        std::vector<u32> hc(h_counts.p, h_counts.p + nhc);
        for (size_t q = 0; q < loc.size(); q++) {
            const size_t k = loc[q];
            const u32 *h   = hc.data() + q * (np + 1);
            if (h[np])
                throw std::runtime_error("a new id could not be computed");
            for (size_t d = 0; d < np; d++)
                if (d != k)
                    cm[k * np + d] = h[d];
        }
    }
    comm_allreduce_host_u64(*this, cm.data(), cm.size(), 1);
    bool any = false;
    for (u64 v : cm)
        any = any || v;
    if (!any)
        return;
    // 2. per local sender that loses objects: ids of those that stay / leave (one scan), the leavers of every
    //    receiver staged as ONE block holding all fields, then the sender compacted — its new size is known from
    //    the counts, no read-back.  (A launch per field and pair, and a scan per pair, cost 30 launches per pair:
    //    28 ms per step for 44 patches of a disc; now 2 per pair + 9 per sender.)
    struct Mig {
        u32 src, dst, count, dst_off = 0;
        DevBuf<f64> blk; ///< [field 0 of all rows | field 1 ...] (local sender: staged rows; remote sender: received)
    };
    std::vector<Mig> migs;
    for (size_t k = 0; k < np; k++)
        for (size_t d = 0; d < np; d++)
            if (cm[k * np + d]) {
                Mig m;
                m.src   = u32(k);
                m.dst   = u32(d);
                m.count = u32(cm[k * np + d]);
                migs.push_back(std::move(m));
            }
    DevBuf<u32> leave_ids, sel_ids;
    for (size_t k = 0, im = 0; k < np; k++) {
        size_t im_end = im;
        u64 out       = 0;
        while (im_end < migs.size() && migs[im_end].src == k)
            out += migs[im_end++].count;
        PatchD &p = patches[k];
        if (out && is_local(p)) {
            const u32 n = p.f.n, kept = u32(n - out);
            flag.ensure(n);
            pos.ensure(n);
            owner_tmp.ensure(n);
            leave_ids.ensure(out);
            flag_equal(s(), n, owners[k].p, u32(k), flag.p);
            exclusive_scan<u8>(s(), flag.p, pos.p, n, scan_tmp, red.p + 5);
            split_ids(s(), n, flag.p, pos.p, owner_tmp.p, leave_ids.p);
            for (size_t q = im; q < im_end; q++) {
                Mig &m = migs[q];
                sel_ids.ensure(m.count);
                select_equal(s(), u32(out), leave_ids.p, owners[k].p, m.dst, sel_ids.p);
                m.blk.ensure(size_t(m.count) * kRowDoubles);
                rows_gather(s(), m.count, sel_ids.p, fields_to_block(p.f, m.blk.p, m.count), 0, 0);
            }
            // 3. compact the sender (order preserving)
            field_tmp.ensure(size_t(kept) * kRowDoubles + 1);
            rows_gather(s(), kept, owner_tmp.p, fields_to_block(p.f, field_tmp.p, kept), 0, 0);
            rows_gather(s(), kept, nullptr, block_to_fields(field_tmp.p, kept, p.f), 0, 0);
            p.f.n = kept;
        }
        im = im_end;
    }
    // 4. append at the destinations: (sender, receiver) ascending
    // (a destination may receive from several senders: reserve once for the sum of the incoming counts)
    std::vector<u64> incoming(np, 0);
    for (auto &m : migs)
        incoming[m.dst] += m.count;
    for (size_t d = 0; d < np; d++) {
        PatchD &dst = patches[d];
        if (!is_local(dst) || !incoming[d])
            continue;
        if (u64(dst.f.n) + incoming[d] > 0xFFFFFFF0ull)
            throw std::overflow_error("patch object count overflows u32 after the reattribution");
        dst.f.reserve(u32(dst.f.n + incoming[d]), s());
    }
    // rows that cross ranks travel as one message per pair (the staged block)
    bool remote = false;
    for (auto &m : migs)
        if (is_local(patches[m.dst]) && !is_local(patches[m.src]))
            m.blk.ensure(size_t(m.count) * kRowDoubles);
    comm_group_start(*this);
    for (auto &m : migs) {
        PatchD &src = patches[m.src];
        PatchD &dst = patches[m.dst];
        const size_t bytes = size_t(m.count) * kRowDoubles * sizeof(f64);
        if (is_local(dst)) {
            m.dst_off = dst.f.n;
            dst.f.n += m.count;
            if (!is_local(src)) {
                comm_recv(*this, m.blk.p, bytes, src.owner);
                remote = true;
            }
        } else if (is_local(src)) {
            comm_send(*this, m.blk.p, bytes, dst.owner);
            remote = true;
        }
    }
    comm_group_end(*this);
    if (remote)
        comm_wait(*this);
    for (auto &m : migs) {
        PatchD &dst = patches[m.dst];
        if (is_local(dst))
            rows_gather(s(), m.count, nullptr, block_to_fields(m.blk.p, m.count, dst.f), 0, m.dst_off);
    }
    SB_CUDA_CHECK(cudaStreamSynchronize(s()));
}

// ---------------------------------------------------------------------------------------------
// ghost zones
// ---------------------------------------------------------------------------------------------
/// SPHUtilities::build_interf_cache + BasicSPHGhostHandler::find_interfaces / gen_id_table_interfaces.
/// The interface plan is a pure function of replicated metadata (ghost_plan.hpp); the metadata
/// (per-patch max h and particle count, per-interface ghost count) is all-reduced over NCCL.
void Model::build_ghost_cache() {
    const size_t np = patches.size();
    const f64 Rkern = cfg.kernel == SHAMB200_KERNEL_M4 ? 2.0 : 3.0;
    if (np > 65536)
        throw std::runtime_error("too many patches (65536 at most)");
    // interactR_patch = max(h) * htol * Rkern  (SPHUtilities.hpp:84-92)
    reset_red();
    for (size_t k = 0; k < np; k++)
        if (is_local(patches[k]) && patches[k].f.n)
            max_reduce(s(), patches[k].f.n, patches[k].f.hpart.p, red.p + 8 + k);
    read_red(8 + int(np));
    std::vector<u64> meta(2 * np, 0); // [ordered max h | count]; remote patches contribute 0
    for (size_t k = 0; k < np; k++)
        if (is_local(patches[k]) && patches[k].f.n) {
            meta[k]      = h_red.p[8 + k];
            meta[np + k] = patches[k].f.n;
        }
    comm_allreduce_host_u64(*this, meta.data(), meta.size(), 1);
    std::vector<f64> interactR(np, std::numeric_limits<f64>::lowest());
    std::vector<u32> pcount(np, 0);
    npart_all = 0;
    for (size_t k = 0; k < np; k++)
        npart_all += meta[np + k];
    for (size_t k = 0; k < np; k++)
        if (meta[np + k]) {
            interactR[k] = ordered_to_f64(meta[k]) * cfg.htol_up_coarse_cycle * Rkern;
            pcount[k]    = u32(meta[np + k]);
        }
    std::vector<IfaceCand> cand
        = plan_interfaces(std::vector<PatchBox>(patches.begin(), patches.end()), box_min, box_max,
                          cfg.bc == SHAMB200_BC_PERIODIC, interactR, pcount);
    // ghost selection, pass A: every local sender classifies its particles against all of its cut boxes at
    // once (64 per launch) — counts per interface, block offsets kept for pass B
    std::vector<u64> counts(cand.size(), 0);
    std::vector<std::vector<size_t>> mine(np);
    for (size_t q = 0; q < cand.size(); q++)
        mine[cand[q].sender].push_back(q);
    std::vector<size_t> hc_off(np, 0); // slots of every local sender in the pinned count read-back
    {
        size_t run = 0;
        for (size_t sd = 0; sd < np; sd++)
            if (is_local(patches[sd]) && patches[sd].f.n && !mine[sd].empty()) {
                hc_off[sd] = run;
                run += 64 * ((mine[sd].size() + 63) / 64);
            }
        h_counts.ensure(run + 64);
    }
    for (size_t sd = 0; sd < np; sd++) {
        PatchD &S = patches[sd];
        if (!is_local(S) || !S.f.n || mine[sd].empty())
            continue;
        const u32 n = S.f.n, nblocks = grid_for(n, 256);
        const size_t nch = (mine[sd].size() + 63) / 64;
        std::vector<f64> hb(nch * 64 * 6, 0.);
        for (size_t j = 0; j < mine[sd].size(); j++)
            for (int d = 0; d < 3; d++) {
                hb[6 * j + d]     = cand[mine[sd][j]].cut_lo[d];
                hb[6 * j + 3 + d] = cand[mine[sd][j]].cut_hi[d];
            }
        field_tmp.ensure(hb.size());
        h2d_small(s(), field_tmp.p, hb.data(), hb.size() * sizeof(f64));
        S.st.gmask.ensure(size_t(n) * nch, 1.1);
        S.st.gblock.ensure(size_t(nblocks) * 64 * nch, 1.1);
        S.st.gtotals.ensure(64 * nch);
        for (size_t ch = 0; ch < nch; ch++) {
            u32 nbox = u32(std::min<size_t>(64, mine[sd].size() - ch * 64));
            ghost_select_count(
                s(), n, S.f.xyz.p, nbox, field_tmp.p + ch * 64 * 6, S.st.gmask.p + ch * n,
                S.st.gblock.p + ch * 64 * size_t(nblocks), S.st.gtotals.p + ch * 64);
        }
        d2h_small(s(), h_counts.p + hc_off[sd], S.st.gtotals.p, 64 * nch * sizeof(u32));
    }
    SB_CUDA_CHECK(cudaStreamSynchronize(s())); // ONE read-back for all senders
    for (size_t sd = 0; sd < np; sd++) {
        PatchD &S = patches[sd];
        if (!is_local(S) || !S.f.n || mine[sd].empty())
            continue;
        const u32 *hc = h_counts.p + hc_off[sd];
        for (size_t j = 0; j < mine[sd].size(); j++)
            counts[mine[sd][j]] = hc[j];
    }
    comm_allreduce_host_u64(*this, counts.data(), counts.size(), 1);
    // pass B: ids of every interface with a local sender (ascending inside an interface)
    std::vector<u32 *> ids_of(cand.size(), nullptr);
    for (size_t sd = 0; sd < np; sd++) {
        PatchD &S = patches[sd];
        if (!is_local(S) || !S.f.n || mine[sd].empty())
            continue;
        const u32 n = S.f.n, nblocks = grid_for(n, 256);
        const size_t nch = (mine[sd].size() + 63) / 64;
        std::vector<u64> base(64 * nch, 0);
        u64 run = 0;
        for (size_t j = 0; j < mine[sd].size(); j++) {
            base[j] = run;
            run += counts[mine[sd][j]];
        }
        S.st.ids_pool.ensure(run, 1.1);
        S.st.gbase.ensure(base.size());
        h2d_small(s(), S.st.gbase.p, base.data(), base.size() * sizeof(u64));
        for (size_t ch = 0; ch < nch; ch++) {
            u32 nbox = u32(std::min<size_t>(64, mine[sd].size() - ch * 64));
            ghost_select_scatter(
                s(), n, S.st.gmask.p + ch * n, nbox, S.st.gblock.p + ch * 64 * size_t(nblocks), S.st.gbase.p + ch * 64,
                S.st.ids_pool.p);
        }
        for (size_t j = 0; j < mine[sd].size(); j++)
            ids_of[mine[sd][j]] = S.st.ids_pool.p + base[j];
    }
    // keep the non-empty interfaces ("prevent sending empty patches")
    ifaces.clear();
    std::vector<u32> ghost_run(np, 0);
    for (size_t q = 0; q < cand.size(); q++) {
        if (!counts[q])
            continue;
        Iface itf;
        itf.sender   = cand[q].sender;
        itf.receiver = cand[q].receiver;
        for (int d = 0; d < 3; d++) {
            itf.offset[d] = cand[q].offset[d];
            itf.ioff[d]   = cand[q].ioff[d];
            itf.cut_lo[d] = cand[q].cut_lo[d];
            itf.cut_hi[d] = cand[q].cut_hi[d];
        }
        itf.count   = u32(counts[q]);
        itf.dst_off = ghost_run[itf.receiver];
        ghost_run[itf.receiver] += itf.count;
        itf.ids = ids_of[q];
        ifaces.push_back(std::move(itf));
    }
    // staging offsets of the interfaces this rank sends to another rank
    send_total = 0;
    for (auto &itf : ifaces) {
        itf.stage_off = send_total;
        if (is_local(patches[itf.sender]) && !is_local(patches[itf.receiver]))
            send_total += itf.count;
    }
    for (size_t k = 0; k < np; k++) {
        patches[k].st.n = patches[k].f.n;
        patches[k].st.m = patches[k].f.n + ghost_run[k];
    }
}

/// BasicSPHGhostHandler::build_comm_merge_positions (BasicSPHGhosts.hpp:294-321,476-514)
void Model::merge_position_ghost() {
    for (auto &p : patches)
        if (is_local(p) && p.f.n)
            p.st.A.ensure(p.st.m, 1.1);
    // the ghosts that leave this rank first: gather into the staging, hand the messages to the communication
    // stream, and build the local part of the merged positions while they travel
    send_stage.ensure(send_total, 1.1);
    {
        Batcher<GhostXyzhJob> bt(s()); // one launch per BATCH_JOBS interfaces (stream_kernels.cuh)
        for (auto &itf : ifaces) {
            PatchD &R = patches[itf.receiver];
            PatchD &S = patches[itf.sender];
            if (is_local(S) && !is_local(R)) // C1: positions + h of the ghosts, 32 B each, straight from the gather
                bt.add({itf.ids, S.f.xyz.p, S.f.hpart.p, send_stage.p + itf.stage_off, itf.offset[0], itf.offset[1],
                        itf.offset[2], itf.count});
        }
        bt.flush();
    }
    comm_group_start(*this);
    for (auto &itf : ifaces) {
        PatchD &R = patches[itf.receiver];
        PatchD &S = patches[itf.sender];
        if (is_local(S) && !is_local(R))
            comm_send(*this, send_stage.p + itf.stage_off, size_t(itf.count) * sizeof(Pack4), R.owner);
        else if (is_local(R) && !is_local(S))
            comm_recv(*this, R.st.A.p + R.st.n + itf.dst_off, size_t(itf.count) * sizeof(Pack4), S.owner);
    }
    comm_group_end(*this);
    {
        Batcher<GhostXyzhJob> bt(s());
        for (auto &p : patches)
            if (is_local(p) && p.f.n)
                bt.add({nullptr, p.f.xyz.p, p.f.hpart.p, p.st.A.p, 0., 0., 0., p.st.n});
        for (auto &itf : ifaces) {
            PatchD &R = patches[itf.receiver];
            PatchD &S = patches[itf.sender];
            if (is_local(R) && is_local(S))
                bt.add({itf.ids, S.f.xyz.p, S.f.hpart.p, R.st.A.p + R.st.n + itf.dst_off, itf.offset[0], itf.offset[1],
                        itf.offset[2], itf.count});
        }
        bt.flush();
    }
    comm_wait(*this);
    if (cfg.keep_step_data)
        for (auto &p : patches)
            if (is_local(p) && p.f.n) {
                p.st.mh_snapshot.ensure(p.st.m);
                unpack_comp(s(), p.st.m, p.st.A.p, 3, 1, p.st.mh_snapshot.p);
            }
}

/// modules::BuildTrees::build_merged_pos_trees (BuildTrees.cpp:26-66)
void Model::build_merged_pos_trees(f64 tol) {
    // all trees up to their leaf counts, ONE synchronisation, then the Karras trees and the boxes
    // ... and the interaction radius of every node (compute_presteps_rint) in the AABB pass
    for (auto &p : patches)
        if (is_local(p) && p.f.n)
            tree_build_begin(
                s(), p.st.tree, reinterpret_cast<const f64 *>(p.st.A.p), 4, p.st.m, nullptr, nullptr, true,
                cfg.tree_reduction_level, cfg.sort_mode);
    SB_CUDA_CHECK(cudaStreamSynchronize(s()));
    for (auto &p : patches)
        if (is_local(p) && p.f.n)
            tree_build_finish(s(), p.st.tree, reinterpret_cast<const f64 *>(p.st.A.p), 4, tol, &p.st.rint);
}
/// Solver::compute_presteps_rint (Solver.cpp:1322-1356)
void Model::compute_presteps_rint(f64 tol) {
    if (!no_fused_rint())
        return; // build_merged_pos_trees left it in p.st.rint (same maxima, same scale)
    for (auto &p : patches)
        if (is_local(p) && p.f.n) {
            p.st.rint.ensure(size_t(p.st.tree.I) + p.st.tree.L);
            tree_field_max(
                s(), p.st.tree, reinterpret_cast<const f64 *>(p.st.A.p) + 3, tol, p.st.rint.p, 4);
        }
}
/// Solver::start_neighbors_cache (Solver.cpp:1364-1386): Morton-sorted storage + the B200 search
void Model::start_neighbors_cache(f64 tol) {
    const f64 Rkern = cfg.kernel == SHAMB200_KERNEL_M4 ? 2.0 : 3.0;
    if (verbose())
        fprintf(stderr, "[shamb200 rank %d] neighbour cache: %zu interfaces\n", rank, ifaces.size());
    K_local         = 0;
    pair_tests_local = 0;
    size_t nloc = 0;
    for (auto &p : patches)
        nloc += is_local(p) && p.f.n;
    if (nloc > 1) { // the searches of all patches, ONE synchronisation, then the (rare) capacity retries
        for (auto &p : patches)
            if (is_local(p) && p.f.n)
                search_prepare_sorted(s(), p.st.tree, p.st.srch, p.st.A.p, p.st.n);
        timer.mark(s(), "neigh_walk"); // walks and lists of the patches alternate: one stage for both
        for (auto &p : patches)
            if (is_local(p) && p.f.n)
                search_enqueue(s(), p.st.tree, p.st.srch, p.st.rint.p, Rkern, tol);
        SB_CUDA_CHECK(cudaStreamSynchronize(s()));
    }
    for (auto &p : patches)
        if (is_local(p) && p.f.n) {
            if (nloc > 1) {
                search_finish(s(), p.st.tree, p.st.srch, Rkern, tol);
            } else {
                search_prepare_sorted(s(), p.st.tree, p.st.srch, p.st.A.p, p.st.n);
                search_build(
                    s(), p.st.tree, p.st.srch, p.st.rint.p, Rkern, tol, [&](const char *name) { timer.mark(s(), name); });
            }
            K_local += p.st.srch.K;
            pair_tests_local += p.st.srch.pair_tests;
            if (verbose() > 1)
                fprintf(stderr, "[shamb200 rank %d]   patch %llu: n %u m %u L %u K %llu attempts %u over %u\n", rank,
                        (unsigned long long) p.id, p.st.n, p.st.m, p.st.tree.L, (unsigned long long) p.st.srch.K,
                        p.st.srch.attempts_last, p.st.srch.over_groups_last);
        }
}

/// Solver::sph_prestep (Solver.cpp:1060-1304)
///
/// List tolerance.  The reference builds its neighbour lists with the radius R h htol (htol_up_coarse_cycle = 1.1)
/// so that h may grow by 10 % inside the step without a new search; a third of the entries of such a list lie
/// outside every kernel support when h barely moves — the usual case.  The fast fp mode builds its lists with a
/// tolerance tol <= htol fitted to what the previous step needed (h_growth_last) and checks afterwards what this
/// step needed: the h iteration reports the largest h_iterate / h_old of any particle and sweep.  If that stayed
/// below tol, every density sum and every later loop saw all the pairs inside its supports (the entries left out
/// contribute exactly nothing) and the step is the one the full lists would have given; if not, h is restored
/// and the sub-cycle is redone with the reference's tolerance (list_fallbacks).  Ghost zones always use htol.
/// The strict mode, keep_step_data (the lists can be exported) and epsilon_h != 1e-6 keep htol.
void Model::sph_prestep() {
    const f64 htol = cfg.htol_up_coarse_cycle;
    f64 tol        = htol;
    if (cfg.fp_mode == SHAMB200_FP_FAST && !cfg.keep_step_data && cfg.epsilon_h == 1e-6) {
        const f64 ov = list_tol_override();
        if (ov < 0)
            tol = list_tol_next > 1. ? std::min(list_tol_next, htol) : htol;
        else if (ov > 1.)
            tol = std::min(ov, htol);
    }
    f64 growth_step = 1.;
    u32 hstep_cnt = 0;
    for (; hstep_cnt < cfg.h_max_subcycles_count; hstep_cnt++) {
        timer.mark(s(), "ghost_cache");
        build_ghost_cache();
        timer.mark(s(), "merge_position_ghost");
        merge_position_ghost();
        timer.mark(s(), "build_trees");
        build_merged_pos_trees(tol);
        timer.mark(s(), "rint");
        compute_presteps_rint(tol);
        timer.mark(s(), "neigh_prepare"); // sorted storage + packed nodes; then "neigh_walk", "neigh_lists"
        start_neighbors_cache(tol);
        timer.mark(s(), "h_iteration");
        if (cfg.gpart_mass == 0)
            throw std::runtime_error(
                "invalid gpart_mass 0, this configuration can not converge.\nPlease set it using either "
                "model.set_particle_mass(pmass) or cfg.set_particle_mass(pmass)");
        for (auto &p : patches)
            if (is_local(p) && p.f.n) {
                PatchStep &st = p.st;
                st.eps.ensure(st.n);
                st.h_old.ensure(st.n);
                fill<f64>(s(), st.eps.p, st.n, 100.);
                SB_CUDA_CHECK(cudaMemcpyAsync(st.h_old.p, p.f.hpart.p, size_t(st.n) * sizeof(f64), cudaMemcpyDeviceToDevice, s()));
            }
        f64 local_max_eps = std::numeric_limits<f64>::max();
        f64 local_min_eps = -1;
        u32 iter_h        = 0;
        // A Newton sweep of particle a reads only positions and its own h (IterateSmoothingLengthDensity.cpp:
        // 52-119), and a particle stops sweeping once its eps <= 1e-6 (hard-coded gate, :75): all sweeps of
        // LoopSmoothingLengthIter run inside ONE launch, each particle iterating on its own, with the same
        // result as the reference's synchronous sweeps.  (epsilon_h != 1e-6 keeps the host loop.)  Ω is
        // evaluated in the same launch with the converged h (ComputeOmega.cpp:36-73).
        const bool fused = cfg.epsilon_h == 1e-6;
        const bool omega_later = omega_in_av_pass(); // fast fp + MM97 / CD10: Ω comes out of av_operators
        if (fused) {
            reset_red();
            for (auto &p : patches)
                if (is_local(p) && p.f.n) {
                    PatchStep &st = p.st;
                    st.omega.ensure(st.n);
                    h_solve(
                        s(), cfg.fp_mode, cfg.kernel, rank_csr_of(st.srch, st.tree), st.srch.SA.p, st.h_old.p,
                        p.f.hpart.p, st.eps.p, st.omega.p, cfg.gpart_mass, cfg.htol_up_coarse_cycle,
                        cfg.htol_up_fine_cycle, cfg.h_iter_per_subcycles, true, !omega_later, red.p);
                }
            read_red(7);
            local_max_eps = ordered_to_f64(h_red.p[0]);
            local_min_eps = ordered_to_f64(h_red.p[1]);
            u32 sweeps    = u32(h_red.p[2]); // slot 2 doubles as the sweep counter during the h iteration
            iter_h        = local_max_eps < cfg.epsilon_h ? (sweeps ? sweeps - 1 : 0) : cfg.h_iter_per_subcycles;
        } else {
            for (; iter_h < cfg.h_iter_per_subcycles; iter_h++) {
                reset_red();
                for (auto &p : patches)
                    if (is_local(p) && p.f.n) {
                        PatchStep &st = p.st;
                        h_solve(
                            s(), cfg.fp_mode, cfg.kernel, rank_csr_of(st.srch, st.tree), st.srch.SA.p, st.h_old.p,
                            p.f.hpart.p, st.eps.p, nullptr, cfg.gpart_mass, cfg.htol_up_coarse_cycle,
                            cfg.htol_up_fine_cycle, 1, true, false, red.p);
                    }
                read_red(2);
                local_max_eps = ordered_to_f64(h_red.p[0]);
                local_min_eps = ordered_to_f64(h_red.p[1]);
                if (local_max_eps < cfg.epsilon_h)
                    break;
            }
        }
        h_iters_last         = iter_h;
        if (verbose())
            fprintf(stderr, "[shamb200 rank %d] h sub-cycle %u: sweeps %u, eps in [%g, %g], K %llu, tests %llu\n", rank,
                    hstep_cnt, iter_h, local_min_eps, local_max_eps, (unsigned long long) K_local,
                    (unsigned long long) pair_tests_local);
        bool should_rerun_gz = local_min_eps < 0;
        bool below_tol       = local_max_eps < cfg.epsilon_h;
        bool converged       = below_tol && !should_rerun_gz;
        // are_all_rank_true (LoopSmoothingLengthIter.cpp:63-64) and the largest h growth of any rank in one
        // min all-reduce (the ordered encoding is monotonic: min of the complement = complement of the max)
        const u64 growth_enc = (fused && cfg.fp_mode == SHAMB200_FP_FAST) ? h_red.p[6] : 0;
        u64 agree[2]         = {converged ? 1ull : 0ull, ~growth_enc};
        comm_allreduce_host_u64(*this, agree, 2, 2);
        converged        = agree[0] != 0;
        const f64 growth = ~agree[1] ? ordered_to_f64(~agree[1]) : 1.;
        if (tol < htol && !(growth <= tol * (1. - 1e-12))) {
            // some h outgrew the lists (a ghost's h grows on the rank that owns it: every rank sees the same
            // maximum): back to the h of before the iteration, once more with the reference's tolerance
            for (auto &p : patches)
                if (is_local(p) && p.f.n)
                    SB_CUDA_CHECK(cudaMemcpyAsync(
                        p.f.hpart.p, p.st.h_old.p, size_t(p.st.n) * sizeof(f64), cudaMemcpyDeviceToDevice, s()));
            if (verbose())
                fprintf(stderr, "[shamb200 rank %d] h grew by %.4f > list tolerance %.4f: sub-cycle redone with %.2f\n",
                        rank, growth, tol, htol);
            tol = htol;
            list_fallbacks++;
            hstep_cnt--;
            continue;
        }
        growth_step = std::max(growth_step, growth);
        if (!converged)
            continue;
        break;
    }
    h_subcycles   = hstep_cnt + 1;
    list_tol_last = tol;
    h_growth_last = growth_step;
    // next step: 2.5 x the growth this one needed + 0.3 %, at least 0.5 %; close to htol there is nothing to gain
    {
        const f64 want = 1. + 2.5 * (growth_step - 1.) + 0.003;
        list_tol_next  = std::max(1.005, want);
        if (!(list_tol_next < 1. + 0.8 * (htol - 1.)))
            list_tol_next = htol;
    }
    if (cfg.epsilon_h != 1e-6 && !omega_in_av_pass()) {
        timer.mark(s(), "omega");
        for (auto &p : patches)
            if (is_local(p) && p.f.n) {
                PatchStep &st = p.st;
                st.omega.ensure(st.n);
                h_solve(
                    s(), cfg.fp_mode, cfg.kernel, rank_csr_of(st.srch, st.tree), st.srch.SA.p, st.h_old.p, p.f.hpart.p,
                    st.eps.p, st.omega.p, cfg.gpart_mass, cfg.htol_up_coarse_cycle, cfg.htol_up_fine_cycle, 0, false,
                    true, red.p);
            }
    }
}

/// Solver::communicate_merge_ghosts_fields (Solver.cpp:1394-1633).  The merged fields are written
/// straight into the Morton-sorted records (slot = inv_map[merged index]).
void Model::communicate_merge_ghosts_fields() {
    const bool has_a = cfg.av == SHAMB200_AV_CD10;
    std::vector<std::pair<const Iface *, size_t>> recv_plan;
    size_t recv_total = 0;
    for (auto &p : patches) {
        if (!is_local(p) || !p.f.n)
            continue;
        PatchStep &st = p.st;
        st.SB.ensure(st.m, 1.1);
        st.SC.ensure(st.m, 1.1);
        if (has_a)
            st.SD.ensure(st.m, 1.1);
    }
    // C2: the ghost fields.  Remote interfaces are staged as [A | B | C | (D)] blocks of Pack4 (one NCCL
    // message per interface) and scattered into the receiver's sorted records on arrival.  The messages are
    // staged and handed to the communication stream first; the records of the local objects and of the local
    // interfaces are packed while they travel.
    const size_t nblk = has_a ? 4 : 3;
    send_stage.ensure(send_total * nblk, 1.1);
    {
        Batcher<GhostXyzhJob> bx(s());
        Batcher<PackFieldsJob> bf(s());
        for (auto &itf : ifaces) {
            PatchD &R = patches[itf.receiver];
            PatchD &S = patches[itf.sender];
            if (is_local(S) && !is_local(R)) {
                Pack4 *sA = send_stage.p + itf.stage_off * nblk;
                Pack4 *sB = sA + itf.count, *sC = sB + itf.count, *sD = has_a ? sC + itf.count : nullptr;
                bx.add({itf.ids, S.f.xyz.p, S.f.hpart.p, sA, itf.offset[0], itf.offset[1], itf.offset[2], itf.count});
                bf.add({itf.ids, nullptr, S.f.hpart.p, S.f.vxyz.p, S.f.uint_.p, S.st.omega.p,
                        has_a ? S.f.axyz.p : nullptr, sA, sB, sC, sD, itf.count});
            } else if (is_local(R) && !is_local(S)) {
                recv_plan.push_back({&itf, recv_total});
                recv_total += size_t(itf.count) * nblk;
            }
        }
        bx.flush(); // positions first: pack_fields overwrites A.d of the same records
        bf.flush();
    }
    recv_stage.ensure(recv_total, 1.1);
    comm_group_start(*this);
    for (auto &itf : ifaces) {
        PatchD &R = patches[itf.receiver];
        PatchD &S = patches[itf.sender];
        if (is_local(S) && !is_local(R))
            comm_send(*this, send_stage.p + itf.stage_off * nblk, size_t(itf.count) * nblk * sizeof(Pack4), R.owner);
    }
    for (auto &rp : recv_plan)
        comm_recv(*this, recv_stage.p + rp.second, size_t(rp.first->count) * nblk * sizeof(Pack4),
                  patches[rp.first->sender].owner);
    comm_group_end(*this);
    {
        Batcher<PackFieldsJob> bf(s());
        for (auto &p : patches) {
            if (!is_local(p) || !p.f.n)
                continue;
            PatchStep &st = p.st;
            bf.add({nullptr, st.srch.inv_map.p, p.f.hpart.p, p.f.vxyz.p, p.f.uint_.p, st.omega.p,
                    has_a ? p.f.axyz.p : nullptr, st.srch.SA.p, st.SB.p, st.SC.p, st.SD.p, st.n});
        }
        for (auto &itf : ifaces) {
            PatchD &R = patches[itf.receiver];
            PatchD &S = patches[itf.sender];
            if (is_local(R) && is_local(S)) {
                u32 o = R.st.n + itf.dst_off;
                bf.add({itf.ids, R.st.srch.inv_map.p + o, S.f.hpart.p, S.f.vxyz.p, S.f.uint_.p, S.st.omega.p,
                        has_a ? S.f.axyz.p : nullptr, R.st.srch.SA.p, R.st.SB.p, R.st.SC.p, has_a ? R.st.SD.p : nullptr,
                        itf.count});
            }
        }
        bf.flush();
    }
    comm_wait(*this);
    {
        Batcher<UnpackGhostJob> bu(s());
        for (auto &rp : recv_plan) {
            const Iface &itf = *rp.first;
            PatchD &R        = patches[itf.receiver];
            u32 o            = R.st.n + itf.dst_off;
            const Pack4 *sA  = recv_stage.p + rp.second;
            const Pack4 *sB = sA + itf.count, *sC = sB + itf.count, *sD = has_a ? sC + itf.count : nullptr;
            bu.add({sA, sB, sC, sD, R.st.srch.inv_map.p + o, R.st.srch.SA.p, R.st.SB.p, R.st.SC.p,
                    has_a ? R.st.SD.p : nullptr, itf.count});
        }
        bu.flush();
    }
}

/// alpha_AV ghost exchange (Solver.cpp:2325-2368).  with_omega (fast fp mode): Ω of every merged object is
/// produced by the operator pass that precedes this exchange (av_operators) and rides along — 16 B per ghost.
void Model::exchange_alpha_ghosts(bool with_omega) {
    // C3: 8 (16) B per ghost; remote interfaces go through compact f64 staging on both sides; the local part is
    // packed while the messages travel
    size_t recv_total = 0;
    for (auto &itf : ifaces)
        if (is_local(patches[itf.receiver]) && !is_local(patches[itf.sender]))
            recv_total += itf.count;
    const size_t nv = with_omega ? 2 : 1; // staging: [alpha of the interface | omega of the interface]
    send_stage_f.ensure(send_total * nv, 1.1);
    recv_stage_f.ensure(recv_total * nv, 1.1);
    for (auto &itf : ifaces) {
        PatchD &R = patches[itf.receiver];
        PatchD &S = patches[itf.sender];
        if (is_local(S) && !is_local(R)) {
            f64 *stg = send_stage_f.p + itf.stage_off * nv;
            gather_field(s(), itf.count, 1, itf.ids, S.st.alpha_updated.p, stg);
            if (with_omega)
                gather_field(s(), itf.count, 1, itf.ids, S.st.omega.p, stg + itf.count);
        }
    }
    size_t roff = 0;
    comm_group_start(*this);
    for (auto &itf : ifaces) {
        PatchD &R = patches[itf.receiver];
        PatchD &S = patches[itf.sender];
        if (is_local(S) && !is_local(R)) {
            comm_send(*this, send_stage_f.p + itf.stage_off * nv, size_t(itf.count) * nv * sizeof(f64), R.owner);
        } else if (is_local(R) && !is_local(S)) {
            comm_recv(*this, recv_stage_f.p + roff, size_t(itf.count) * nv * sizeof(f64), S.owner);
            roff += size_t(itf.count) * nv;
        }
    }
    comm_group_end(*this);
    {
        Batcher<PackAlphaJob> ba(s());
        for (auto &p : patches)
            if (is_local(p) && p.f.n)
                ba.add({nullptr, p.st.srch.inv_map.p, p.st.alpha_updated.p, with_omega ? p.st.omega.p : nullptr,
                        p.st.SC.p, p.st.n});
        for (auto &itf : ifaces) {
            PatchD &R = patches[itf.receiver];
            PatchD &S = patches[itf.sender];
            if (is_local(R) && is_local(S))
                ba.add({itf.ids, R.st.srch.inv_map.p + R.st.n + itf.dst_off, S.st.alpha_updated.p,
                        with_omega ? S.st.omega.p : nullptr, R.st.SC.p, itf.count});
        }
        ba.flush();
    }
    comm_wait(*this);
    {
        Batcher<PackAlphaJob> ba(s());
        roff = 0;
        for (auto &itf : ifaces) {
            PatchD &R = patches[itf.receiver];
            if (is_local(R) && !is_local(patches[itf.sender])) {
                const f64 *stg = recv_stage_f.p + roff;
                ba.add({nullptr, R.st.srch.inv_map.p + R.st.n + itf.dst_off, stg, with_omega ? stg + itf.count : nullptr,
                        R.st.SC.p, itf.count});
                roff += size_t(itf.count) * nv;
            }
        }
        ba.flush();
    }
}

// ---------------------------------------------------------------------------------------------
// the step
// ---------------------------------------------------------------------------------------------
void Model::evolve_once() {
    auto wall0 = std::chrono::steady_clock::now();
    SB_CUDA_CHECK(cudaSetDevice(ctx->device));
    timer.begin_step();
    const f64 t_current = time;
    const f64 dt_       = dt;
    const bool has_alpha  = cfg.av == SHAMB200_AV_MM97 || cfg.av == SHAMB200_AV_CD10;
    const bool has_curl   = cfg.av == SHAMB200_AV_CD10;
    const bool has_dtdivv = cfg.av == SHAMB200_AV_CD10;
    const bool has_cs_field = has_alpha || cfg.eos == SHAMB200_EOS_LOCALLY_ISOTHERMAL_LP07;
    red.ensure(red_slots(*this));
    h_red.ensure(red_slots(*this));

    // Solver.cpp:1970-1976: the scheduler step (split / merge / load balancing) opens the step
    if (scheduler_freq && step_count % scheduler_freq == 0 && !pipe.active) {
        timer.mark(s(), "scheduler");
        scheduler_step(true, true);
    }
    timer.mark(s(), "predictor");
    point_mass_accrete_particles();
    // host-resident patch data (evolve_once_host): copies overlap the kernels
    const bool piped_in = pipe.active && pipe.defer_in2, piped = pipe.active && pipe.early_out;
    // periodic box and nothing that reads the positions between the drift and the position boundary (kill
    // spheres, point-mass force): the wrap of apply_position_boundary rides in the drift kernel (same arithmetic)
    wrapped_in_drift = cfg.bc == SHAMB200_BC_PERIODIC && cfg.n_kill_spheres == 0 && !cfg.has_point_mass;
    const f64 *wmin = wrapped_in_drift ? box_min : nullptr, *wmax = wrapped_in_drift ? box_max : nullptr;
    for (auto &p : patches)
        if (is_local(p) && p.f.n) {
            if (piped_in) // uint / duint are still uploading: drift the positions now, u after the prestep
                leapfrog_predictor_pos(s(), p.f.n, dt_, p.f.xyz.p, p.f.vxyz.p, p.f.axyz.p, wmin, wmax);
            else
                leapfrog_predictor(
                    s(), p.f.n, dt_, p.f.xyz.p, p.f.vxyz.p, p.f.axyz.p, p.f.uint_.p, p.f.duint.p, wmin, wmax);
        }
    kill_particles();
    compute_ext_forces_indep_v();
    timer.mark(s(), "position_boundary");
    apply_position_boundary();
    // (the global object count comes out of the ghost-zone metadata all-reduce of the prestep)

    // Solver.cpp:2043-2048
    if (cfg.enable_particle_reordering && cfg.particle_reordering_step_freq
        && step_count % cfg.particle_reordering_step_freq == 0) {
        if (pipe.active && pipe.defer_in2) // host-resident step: every input field must have arrived
            SB_CUDA_CHECK(cudaStreamWaitEvent(s(), pipe.ev_in2, 0));
        reorder_particles();
    }

    if (piped) { // are consecutive ids neighbours in space?  (pipe_slices; the search synchronises before it is read)
        const PatchD &p = patches[pipe.ip];
        pipe.far_dev.ensure(1);
        pipe.far_host.ensure(1);
        count_far_successors(s(), p.f.n, p.f.xyz.p, p.f.hpart.p, pipe.far_dev.p);
        d2h_small(s(), pipe.far_host.p, pipe.far_dev.p, sizeof(u64));
    }
    if (piped) { // positions (and the external acceleration) are final once the boundary has been applied: they
        static const char *const first[] = {"xyz", "axyz_ext"}; // go back while the tree and the cache are built
        pipe_download(first, 2);
    }
    sph_prestep();
    if (piped) {
        static const char *const early[] = {"hpart"};
        pipe_download(early, 1); // final since the h iteration: goes back during the CD10 operators
    }
    if (piped_in) {
        SB_CUDA_CHECK(cudaStreamWaitEvent(s(), pipe.ev_in2, 0));
        for (auto &p : patches)
            if (is_local(p) && p.f.n)
                leapfrog_predictor_u(s(), p.f.n, dt_, p.f.uint_.p, p.f.duint.p);
    }

    SphParams sp{cfg.gpart_mass, cfg.alpha_u, cfg.alpha_AV, cfg.beta_AV};
    f64 next_cfl              = 0;
    u32 corrector_iter_cnt    = 0;
    bool need_rerun_corrector = false;
    do {
        if (corrector_iter_cnt == 50)
            throw std::runtime_error(
                "the corrector has made over 50 loops, either their is a bug, either you are using a dt that "
                "is too large");
        timer.mark(s(), "ghost_fields");
        communicate_merge_ghosts_fields();
        for (auto &p : patches)
            if (is_local(p) && p.f.n && has_alpha) {
                p.st.alpha_updated.ensure(p.st.n);
                SB_CUDA_CHECK(cudaMemcpyAsync(
                    p.st.alpha_updated.p, p.f.alpha_AV.p, size_t(p.st.n) * sizeof(f64), cudaMemcpyDeviceToDevice, s()));
            }
        timer.mark(s(), "divv_curlv_dtdivv");
        // host-resident step: the two heavy loops (and the corrector behind the force loop) run over id ranges,
        // the outputs of a finished range travel to the host while the next one is computed (pipe_slices)
        const std::vector<std::pair<u32, u32>> slices =
            piped && corrector_iter_cnt == 0 ? pipe_slices() : std::vector<std::pair<u32, u32>>{};
        const bool sliced = slices.size() > 1;
        static const char *const ops_out[] = {"divv", "curlv", "dtdivv"};
        if (has_alpha)
            for (auto &p : patches) {
                if (!is_local(p) || !p.f.n)
                    continue;
                PatchStep &st = p.st;
                st.omega.ensure(st.n);
                auto run = [&](const RankCsr &c) {
                    av_operators(
                        s(), cfg.fp_mode, cfg.kernel, c, st.srch.SA.p, st.SB.p, st.SC.p, st.SD.p, cfg.gpart_mass,
                        has_curl, has_dtdivv, cfg.combined_dtdiv_divcurlv_compute != 0, p.f.divv.p, p.f.curlv.p,
                        p.f.dtdivv.p, omega_in_av_pass() ? st.omega.p : nullptr);
                };
                const RankCsr csr = rank_csr_of(st.srch, st.tree);
                if (sliced) // (one local patch: pipe.early_out)
                    for (auto &sl : slices) {
                        run(csr.ids(sl.first, sl.second));
                        pipe_download(ops_out, 3, sl.first, sl.second);
                    }
                else
                    run(csr);
            }
        if (piped && corrector_iter_cnt == 0 && has_alpha && !sliced) // back during the EOS / forces
            pipe_download(ops_out, 3);
        timer.mark(s(), "av_eos");
        if (has_alpha) {
            for (auto &p : patches)
                if (is_local(p) && p.f.n)
                    update_av(
                        s(), cfg.av, p.st.n, dt_, cfg.sigma_decay, cfg.alpha_min, cfg.alpha_max, p.f.divv.p,
                        p.f.curlv.p, p.f.dtdivv.p, p.f.soundspeed.p, p.f.hpart.p, p.f.alpha_AV.p, p.st.alpha_updated.p);
            exchange_alpha_ghosts(omega_in_av_pass());
        }
        for (auto &p : patches)
            if (is_local(p) && p.f.n)
                compute_eos(
                    s(), cfg.kernel, cfg.eos, p.st.srch.SA.p, p.st.SB.p, p.st.SC.p, p.st.m, cfg.gpart_mass, cfg.gamma,
                    cfg.cs0, cfg.eos_q, cfg.eos_r0);
        // adiabatic EOS: the force loop reads a 16-byte third record and recomputes the neighbour's rho and P
        // (SHAMB200_SF16=0 keeps the 32-byte record: tuning runs)
        const char *sf16_env = getenv("SHAMB200_SF16");
        const bool sf16 = cfg.fp_mode == SHAMB200_FP_FAST && cfg.eos == SHAMB200_EOS_ADIABATIC
                          && !(sf16_env && atoi(sf16_env) == 0);
        if (cfg.fp_mode == SHAMB200_FP_FAST)
            for (auto &p : patches)
                if (is_local(p) && p.f.n) {
                    p.st.SE.ensure(p.st.m, 1.1);
                    p.st.SF.ensure(p.st.m, 1.1);
                    if (sf16)
                        p.st.SG.ensure(p.st.m, 1.1);
                    derive_fast(
                        s(), cfg.kernel, cfg.av, p.st.m, p.st.srch.SA.p, p.st.SB.p, p.st.SC.p, cfg.gpart_mass,
                        cfg.alpha_AV, p.st.SE.p, p.st.SF.p, sf16 ? p.st.SG.p : nullptr);
                }
        if (piped && corrector_iter_cnt == 0) {
            static const char *const mid[] = {"alpha_AV@updated"};
            pipe_download(mid, has_alpha ? 1 : 0); // goes back during the force loop
        }
        timer.mark(s(), "forces");
        // forces, v_sig and the CFL dt come out of one pass over the neighbour lists.  The CFL uses the cfl
        // multiplier in force at launch: if the corrector test below halves it, the whole pass is redone.
        const f64 C_cour  = cfg.cfl_cour * cfl_multiplier;
        const f64 C_force = cfg.cfl_force * cfl_multiplier;
        reset_red();
        step_sc.ensure(16);
        h_step_sc.ensure(16);
        SB_CUDA_CHECK(cudaMemsetAsync(step_sc.p, 0, 16 * sizeof(f64), s()));
        for (auto &p : patches) {
            if (!is_local(p) || !p.f.n)
                continue;
            PatchStep &st = p.st;
            st.a_old.ensure(size_t(st.n) * 3);
            st.du_old.ensure(st.n);
            st.vsig.ensure(st.n);
            st.cfl_dt.ensure(st.n);
            SB_CUDA_CHECK(cudaMemcpyAsync(st.a_old.p, p.f.axyz.p, size_t(st.n) * 3 * sizeof(f64), cudaMemcpyDeviceToDevice, s()));
            SB_CUDA_CHECK(cudaMemcpyAsync(st.du_old.p, p.f.duint.p, size_t(st.n) * sizeof(f64), cudaMemcpyDeviceToDevice, s()));
            SphParams spf = sp;
            if (sf16) {
                spf.adiabatic_gm1 = cfg.gamma - 1;
                spf.SG            = st.SG.p;
            }
            auto run = [&](const RankCsr &c) {
                force_cfl(
                    s(), cfg.fp_mode, cfg.kernel, cfg.av, c, st.srch.SA.p, st.SB.p, st.SC.p, st.SE.p, st.SF.p, spf,
                    p.f.axyz_ext.p, p.f.axyz.p, p.f.duint.p, C_cour, C_force, st.vsig.p, st.cfl_dt.p, red.p + 4);
            };
            const RankCsr csr = rank_csr_of(st.srch, st.tree);
            if (sliced) { // forces, then the corrector of the same id range, then its four fields go home
                static const char *const done[] = {"axyz", "duint", "vxyz", "uint"};
                for (auto &sl : slices) {
                    const u32 i0 = sl.first, n = sl.second;
                    run(csr.ids(i0, n));
                    leapfrog_corrector(
                        s(), n, dt_ / 2, p.f.vxyz.p + 3 * size_t(i0), p.f.axyz.p + 3 * size_t(i0),
                        st.a_old.p + 3 * size_t(i0), p.f.uint_.p + i0, p.f.duint.p + i0, st.du_old.p + i0, red.p + 2,
                        reinterpret_cast<f64 *>(red.p + 3), step_sc.p + 1);
                    pipe_download(done, 4, i0, n);
                }
            } else {
                run(csr);
            }
        }
        if (piped && corrector_iter_cnt == 0 && !sliced) { // the accelerations are final before the corrector uses them
            static const char *const acc[] = {"axyz", "duint"};
            pipe_download(acc, 2);
        }
        timer.mark(s(), "corrector");
        if (!sliced)
            for (auto &p : patches)
                if (is_local(p) && p.f.n)
                    leapfrog_corrector(
                        s(), p.st.n, dt_ / 2, p.f.vxyz.p, p.f.axyz.p, p.st.a_old.p, p.f.uint_.p, p.f.duint.p,
                        p.st.du_old.p, red.p + 2, reinterpret_cast<f64 *>(red.p + 3), step_sc.p + 1);
        // C7 + C8 (Solver.cpp:2587, :2597, :3119) and the sums of ConservativeCheck, fused: Σ v² and the
        // conservation sums in one sum all-reduce, max eps_v² and -min dt in one max all-reduce, both queued on
        // the stream, then ONE copy and ONE synchronisation (the dt is only used if the corrector holds)
        step_scalars(s(), red.p, step_sc.p);
        comm_allreduce_f64(*this, step_sc.p, 9, 0);
        comm_allreduce_f64(*this, step_sc.p + 9, 2, 1);
        d2h_small(s(), h_step_sc.p, step_sc.p, 11 * sizeof(f64));
        SB_CUDA_CHECK(cudaStreamSynchronize(s()));
        f64 rank_veps_v = std::sqrt(h_step_sc.p[9]);
        f64 sum_vsq     = h_step_sc.p[0];
        for (int k = 0; k < 8; k++) // ConservativeCheck multiplies by the particle mass (ConservativeCheck.cpp:62-188)
            conservation[k] = cfg.gpart_mass * h_step_sc.p[1 + k];
        f64 vmean_sq   = sum_vsq / f64(npart_all);
        f64 vmean      = std::sqrt(vmean_sq);
        f64 rank_eps_v = rank_veps_v / vmean;
        if (vmean <= 0)
            rank_eps_v = 0;
        eps_v = rank_eps_v;
        if (eps_v > 1e-2) {
            need_rerun_corrector = true;
            cfl_multiplier       = cfl_multiplier / 2;
        } else {
            need_rerun_corrector = false;
        }
        if (!need_rerun_corrector) {
            for (auto &p : patches) {
                if (!is_local(p) || !p.f.n)
                    continue;
                PatchStep &st = p.st;
                if (has_alpha)
                    SB_CUDA_CHECK(cudaMemcpyAsync(
                        p.f.alpha_AV.p, st.alpha_updated.p, size_t(st.n) * sizeof(f64), cudaMemcpyDeviceToDevice, s()));
                if (has_cs_field)
                    unpack_cs(s(), st.n, st.SC.p, p.f.soundspeed.p, st.srch.inv_map.p);
            }
            next_cfl = -h_step_sc.p[10];
        }
        corrector_iter_cnt++;
    } while (need_rerun_corrector);
    corrector_iter = corrector_iter_cnt;
    if (piped) {
        // what left early is final unless the corrector pass was repeated (then it was recomputed: send it again);
        // a configuration without the CD10 fields still returns the (untouched) arrays: every field comes back
        static const char *const tail[] = {"soundspeed", "vxyz", "uint"};
        pipe_download(tail, (pipe.sliced_out && corrector_iter_cnt == 1) ? 1 : 3); // sliced: v, u left with the ranges
        if (corrector_iter_cnt > 1) {
            static const char *const again[] = {"axyz", "duint"};
            pipe_download(again, 2);
        }
        if (corrector_iter_cnt > 1 || !has_alpha) {
            static const char *const cd10[] = {"divv", "curlv", "dtdivv", "alpha_AV"};
            pipe_download(cd10, 4);
        }
    }
    timer.end_step(s());
    SB_CUDA_CHECK(cudaStreamSynchronize(s()));

    dt   = next_cfl;
    time = t_current + dt_;
    f64 stiff      = cfg.cfl_multiplier_stiffness;
    cfl_multiplier = (cfl_multiplier * stiff + 1.) / (stiff + 1.);
    step_count++;
    t_step         = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
}

// ---------------------------------------------------------------------------------------------
// host-resident patch data (shamb200_model_evolve_once_host)
// ---------------------------------------------------------------------------------------------
static f64 *host_field(const shamb200_host_patchdata &h, const std::string &nm) {
    if (nm == "xyz") return h.xyz;
    if (nm == "vxyz") return h.vxyz;
    if (nm == "axyz") return h.axyz;
    if (nm == "axyz_ext") return h.axyz_ext;
    if (nm == "hpart") return h.hpart;
    if (nm == "uint") return h.uint_;
    if (nm == "duint") return h.duint;
    if (nm == "alpha_AV") return h.alpha_AV;
    if (nm == "divv") return h.divv;
    if (nm == "dtdivv") return h.dtdivv;
    if (nm == "curlv") return h.curlv;
    if (nm == "soundspeed") return h.soundspeed;
    return nullptr;
}

/// device -> host copies of the named fields on the download stream, ordered after the work queued so
/// far on the main stream ("alpha_AV@updated": the new alpha before it is committed to the field)
/// first / n: the objects [first, first + n) only (n = 0xffffffff: the whole field)
void Model::pipe_download(const char *const *names, int count, u32 first, u32 n) {
    if (!pipe.out || count <= 0)
        return;
    PatchD &p = patches[pipe.ip];
    const bool part = n != 0xffffffffu;
    if (part && (u64(first) + n > p.f.n))
        throw std::logic_error("pipe_download: range past the end of the patch");
    SB_CUDA_CHECK(cudaEventRecord(pipe.ev_stage, s()));
    SB_CUDA_CHECK(cudaStreamWaitEvent(pipe.d2h, pipe.ev_stage, 0));
    for (int k = 0; k < count; k++) {
        std::string nm(names[k]);
        const f64 *src = nullptr;
        int nvar       = 1;
        if (nm == "alpha_AV@updated") {
            nm  = "alpha_AV";
            src = p.st.alpha_updated.p;
        } else {
            for (auto &r : p.f.all())
                if (nm == r.name) {
                    src  = r.buf->p;
                    nvar = r.nvar;
                }
        }
        f64 *dst = host_field(*pipe.out, nm);
        if (!dst || !src || !p.f.n)
            continue;
        if (p.f.n > pipe.out_cap)
            throw std::length_error("evolve_once_host: the patch holds more objects than out->n (capacity)");
        const size_t o0 = part ? size_t(first) * nvar : 0;
        size_t bytes    = size_t(part ? n : p.f.n) * nvar * sizeof(f64);
        if (!bytes)
            continue;
        SB_CUDA_CHECK(cudaMemcpyAsync(dst + o0, src + o0, bytes, cudaMemcpyDeviceToHost, pipe.d2h));
        pipe.bytes_d2h += bytes;
    }
}

/// id ranges of the host-resident step's sliced loops: SHAMB200_HOST_SLICES (default 8) ranges of at least
/// SHAMB200_HOST_SLICE_MIN (default 2^18) objects, or none — a patch whose consecutive ids are not neighbours in
/// space (more than 1 % of the objects farther than 8 h from their successor: count_far_successors) keeps the
/// slot-ordered launches: neighbouring lanes must share neighbours for the gathers to hit L1.  Patch data that
/// went through ParticleReordering follows a Morton curve (of the patch box, not of the merged tree: slots and
/// ids differ, but either order is local) and takes the ranges
std::vector<std::pair<u32, u32>> Model::pipe_slices() {
    std::vector<std::pair<u32, u32>> out;
    pipe.sliced_out = false;
    if (!pipe.active || !pipe.early_out || !pipe.out)
        return out;
    const char *e_k = getenv("SHAMB200_HOST_SLICES"), *e_min = getenv("SHAMB200_HOST_SLICE_MIN"); // (tests)
    const u32 want  = e_k ? u32(std::max(1, atoi(e_k))) : 8u;
    const u32 least = e_min ? u32(std::max(1, atoi(e_min))) : (1u << 18);
    const PatchD &p = patches[pipe.ip];
    const u32 n     = p.f.n;
    const u32 k     = std::min<u32>(want, std::max<u32>(1u, n / least));
    const char *e_far = getenv("SHAMB200_HOST_SLICE_FAR_PCT"); // (tests: 100 takes the ranges whatever the order)
    const u64 far_pct = e_far ? u64(std::max(0, atoi(e_far))) : 1u;
    if (k < 2 || pipe.far_host.p[0] * 100 > u64(n) * far_pct)
        return out;
    for (u32 q = 0; q < k; q++) {
        const u32 i0 = u32(u64(n) * q / k), i1 = u32(u64(n) * (q + 1) / k);
        out.emplace_back(i0, i1 - i0);
    }
    pipe.sliced_out = true;
    pipe.nslices    = k;
    return out;
}

void Model::evolve_once_host(u32 ip, const shamb200_host_patchdata *in, shamb200_host_patchdata *out) {
    SB_CUDA_CHECK(cudaSetDevice(ctx->device));
    PatchD &p = patches.at(ip);
    if (!is_local(p))
        throw std::invalid_argument("patch is not local");
    if (!pipe.h2d) {
        SB_CUDA_CHECK(cudaStreamCreateWithFlags(&pipe.h2d, cudaStreamNonBlocking));
        SB_CUDA_CHECK(cudaStreamCreateWithFlags(&pipe.d2h, cudaStreamNonBlocking));
        SB_CUDA_CHECK(cudaEventCreateWithFlags(&pipe.ev_in1, cudaEventDisableTiming));
        SB_CUDA_CHECK(cudaEventCreateWithFlags(&pipe.ev_in2, cudaEventDisableTiming));
        SB_CUDA_CHECK(cudaEventCreateWithFlags(&pipe.ev_stage, cudaEventDisableTiming));
    }
    u32 nlocal = 0;
    for (auto &q : patches)
        nlocal += is_local(q) ? 1 : 0;
    const bool has_alpha = cfg.av == SHAMB200_AV_MM97 || cfg.av == SHAMB200_AV_CD10;
    // outputs can leave early once the object count and order are final (after the prestep); the second
    // input group may only arrive late when nothing reorders the patch before (kill, accretion, migration)
    pipe.early_out = nlocal == 1;
    pipe.defer_in2 = pipe.early_out && world == 1 && !cfg.has_point_mass && cfg.n_kill_spheres == 0
                     && cfg.bc == SHAMB200_BC_PERIODIC;
    pipe.ip = ip, pipe.in = in, pipe.out = out;
    pipe.out_cap = out ? out->n : 0;
    pipe.bytes_h2d = pipe.bytes_d2h = 0;
    pipe.sliced_out = false;
    pipe.nslices    = 0;
    SB_CUDA_CHECK(cudaStreamSynchronize(s())); // nothing of a previous call may still read the fields
    if (in && in->n != p.f.n) {                // the host owns the data: the patch takes its size
        if (in->n > 0xFFFFFFF0ull)
            throw std::invalid_argument("patch too large");
        p.f.reserve(u32(in->n), s());
        p.f.n = u32(in->n);
        SB_CUDA_CHECK(cudaStreamSynchronize(s()));
    }
    auto upload = [&](const char *const *names, int count, cudaEvent_t done) {
        for (int k = 0; in && k < count; k++) {
            const f64 *src = host_field(*in, names[k]);
            if (!src || !p.f.n)
                continue;
            for (auto &r : p.f.all())
                if (std::string(names[k]) == r.name) {
                    size_t bytes = size_t(p.f.n) * r.nvar * sizeof(f64);
                    SB_CUDA_CHECK(cudaMemcpyAsync(r.buf->p, src, bytes, cudaMemcpyHostToDevice, pipe.h2d));
                    pipe.bytes_h2d += bytes;
                }
        }
        SB_CUDA_CHECK(cudaEventRecord(done, pipe.h2d));
    };
    static const char *const in1[] = {"xyz", "vxyz", "axyz", "hpart"}; // what the drift and the search read
    // soundspeed is an INPUT of the AV switch (UpdateViscosity.cpp:52-222 reads the previous step's value)
    static const char *const in2[] = {"uint", "duint", "alpha_AV", "soundspeed"};
    upload(in1, 4, pipe.ev_in1);
    upload(in2, has_alpha ? 4 : 2, pipe.ev_in2);
    SB_CUDA_CHECK(cudaStreamWaitEvent(s(), pipe.ev_in1, 0));
    if (!pipe.defer_in2)
        SB_CUDA_CHECK(cudaStreamWaitEvent(s(), pipe.ev_in2, 0));
    pipe.active = true;
    try {
        evolve_once();
    } catch (...) {
        pipe.active = false;
        cudaStreamSynchronize(pipe.h2d);
        cudaStreamSynchronize(pipe.d2h);
        throw;
    }
    pipe.active = false;
    if (!pipe.early_out) { // several local patches: all at the end
        static const char *const every[] = {"xyz", "vxyz", "axyz", "axyz_ext", "hpart", "uint", "duint",
                                            "alpha_AV", "divv", "dtdivv", "curlv", "soundspeed"};
        pipe_download(every, 12);
    }
    SB_CUDA_CHECK(cudaStreamSynchronize(pipe.d2h));
    if (out)
        out->n = p.f.n;
}

} // namespace sb
