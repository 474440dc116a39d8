// -------------------------------------------------------------------------------------------
// sph_step.hpp — CPU restatement of shammodels::sph::Solver::evolve_once (oracle, test-only).
//
// TEST INFRASTRUCTURE ONLY — see the header of shamrock_oracle.hpp for the rules and the
// parity status.  Every function cites the reference lines it follows (paths relative to
// /root/reference/src).
//
// Supported configuration subset: kernels M4/M6; EOS adiabatic / isothermal /
// locally-isothermal LP07; AV constant / MM97 / CD10 / constant-disc; BC free / periodic;
// external force: central point mass (with accretion radius); kill spheres; density-based h.
// Patches: a grid of patches on the 2^21 integer grid; split_patch / merge_patches restate the scheduler's
// patch-list operations (ids, order, data partition); no load balancing (one process).
// -------------------------------------------------------------------------------------------
#pragma once
#include "shamrock_oracle.hpp"
#include <map>
#include <memory>
#include <string>

namespace oracle {

enum KernelId { KERNEL_M4 = 0, KERNEL_M6 = 1 };
enum EosId { EOS_ADIABATIC = 0, EOS_ISOTHERMAL = 1, EOS_LOCALLY_ISOTHERMAL_LP07 = 2 };
enum AvId { AV_NONE = 0, AV_CONSTANT = 1, AV_MM97 = 2, AV_CD10 = 3, AV_CONSTANT_DISC = 4 };
enum BcId { BC_FREE = 0, BC_PERIODIC = 1 };

/// Mirrors the fields of shammodels::sph::SolverConfig that the hot path reads
/// (shammodels/sph/include/shammodels/sph/SolverConfig.hpp:443-630, config/AVConfig.hpp)
struct SolverConfig {
    int kernel       = KERNEL_M4;
    f64 gpart_mass   = 0;
    int eos          = EOS_ADIABATIC;
    f64 gamma        = 5. / 3.;
    f64 cs0          = 1;
    f64 eos_q        = 0;
    f64 eos_r0       = 1;
    int av           = AV_CONSTANT;
    f64 alpha_u      = 1;
    f64 alpha_AV     = 1;
    f64 beta_AV      = 2;
    f64 alpha_min    = 0.1;
    f64 alpha_max    = 1;
    f64 sigma_decay  = 0.1;
    int bc           = BC_FREE;
    f64 cfl_cour     = 0.3;
    f64 cfl_force    = 0.25;
    f64 cfl_multiplier_stiffness = 2;
    f64 htol_up_coarse_cycle     = 1.1;
    f64 htol_up_fine_cycle       = 1.1;
    f64 epsilon_h                = 1e-6;
    u32 h_iter_per_subcycles     = 50;
    u32 h_max_subcycles_count    = 100;
    u32 tree_reduction_level     = 3;
    int use_two_stage_search     = 1;
    int combined_dtdiv_divcurlv_compute = 0;
    // external forces (ext_force_config) : central point mass at the origin
    int has_point_mass = 0;
    f64 pm_mass        = 0;
    f64 pm_racc        = 0;
    f64 constant_G     = 1;
    // Morton reordering of the patch data (SolverConfig.hpp:621-630)
    int enable_particle_reordering    = 0;
    u64 particle_reordering_step_freq = 1000;
    // particle killing (kill spheres)
    int n_kill_spheres = 0;
    f64 kill_center[4][3];
    f64 kill_radius[4];

    bool has_alphaAV() const { return av == AV_MM97 || av == AV_CD10; }
    bool has_divv() const { return has_alphaAV(); }
    bool has_curlv() const { return av == AV_CD10; }
    bool has_dtdivv() const { return av == AV_CD10; }
    bool has_axyz_in_ghost() const { return has_dtdivv(); }
    bool has_soundspeed_field() const {
        return has_alphaAV() || eos == EOS_LOCALLY_ISOTHERMAL_LP07;
    }
};

/// Main patch data layout.  ref: shammodels/sph/src/SolverConfig.cpp:24-121
struct PatchData {
    u32 n = 0;
    std::vector<f64> xyz, vxyz, axyz, axyz_ext;      // 3n, packed
    std::vector<f64> hpart, uint, duint;             // n
    std::vector<f64> alpha_AV, divv, dtdivv, soundspeed; // n (when the config has them)
    std::vector<f64> curlv;                          // 3n

    void resize(u32 nn) {
        n = nn;
        for (auto *v : {&xyz, &vxyz, &axyz, &axyz_ext, &curlv})
            v->resize(size_t(3) * nn, 0.);
        for (auto *v : {&hpart, &uint, &duint, &alpha_AV, &divv, &dtdivv, &soundspeed})
            v->resize(nn, 0.);
    }
    std::vector<std::pair<std::vector<f64> *, int>> all_fields() {
        return {{&xyz, 3},   {&vxyz, 3},  {&axyz, 3},     {&axyz_ext, 3}, {&hpart, 1},
                {&uint, 1},  {&duint, 1}, {&alpha_AV, 1}, {&divv, 1},     {&dtdivv, 1},
                {&curlv, 3}, {&soundspeed, 1}};
    }
    /// PatchDataLayer::keep_ids — keeps the given ids, in the given order
    void keep_ids(const std::vector<u32> &ids) {
        for (auto [f, nv] : all_fields()) {
            std::vector<f64> nf(ids.size() * nv);
            for (size_t k = 0; k < ids.size(); k++)
                for (int c = 0; c < nv; c++)
                    nf[k * nv + c] = (*f)[size_t(ids[k]) * nv + c];
            *f = std::move(nf);
        }
        n = u32(ids.size());
    }
    /// append the subset `ids` of `src` (append_subset_to)
    void append_subset_from(PatchData &src, const std::vector<u32> &ids) {
        auto fs = all_fields();
        auto fo = src.all_fields();
        for (size_t q = 0; q < fs.size(); q++) {
            int nv = fs[q].second;
            for (u32 id : ids)
                for (int c = 0; c < nv; c++)
                    fs[q].first->push_back((*fo[q].first)[size_t(id) * nv + c]);
        }
        n += u32(ids.size());
    }
};

/// A patch on the 2^21 integer grid.  ref: shamrock/include/shamrock/patch/Patch.hpp:63-72,
/// PatchCoord.hpp:135-139 (range = [coord_min, coord_max + 1))
struct Patch {
    u64 id;
    u64 coord_min[3], coord_max[3];
    PatchData pdat;
};

struct Interface {
    u64 sender, receiver;
    f64 offset[3];
    i32 ioff[3];
    f64 cut_lo[3], cut_hi[3];
    std::vector<u32> ids; // indices in the sender patch
};

/// Everything the prestep builds for one patch (kept for parity inspection)
struct PatchStep {
    u32 n = 0, m = 0;               // real, real+ghost
    std::vector<f64> mxyz, mh;      // merged_xyzh (3m, m)
    Tree<u32> tree;
    std::vector<f64> rint;          // I+L
    ObjectCache cache;
    std::vector<f64> omega;         // n
    // merged ghost fields (communicate_merge_ghosts_fields)
    std::vector<f64> g_h, g_u, g_v, g_a, g_omega, g_cs, g_alpha; // m, m, 3m, 3m, m, m, m
    std::vector<f64> pressure, soundspeed;                        // m
    std::vector<f64> alpha_updated;                               // n
    std::vector<f64> vsig, cfl_dt;                                // n
};

struct StepLog {
    u32 h_subcycles    = 0;
    u32 h_iters_last   = 0;
    u32 corrector_iter = 0;
    f64 eps_v          = 0;
    f64 next_dt        = 0;
    u64 npart          = 0;
};

template<class K>
struct SolverT;

/// Type-erased front (kernel chosen at run time)
struct Solver {
    SolverConfig cfg;
    f64 box_min[3] = {0, 0, 0}, box_max[3] = {1, 1, 1};
    std::vector<Patch> patches; // sorted by id
    f64 time = 0, dt = 0, cfl_multiplier = 1e-2; // ref: Solver.hpp:131-147
    std::map<u64, PatchStep> step; // last step's intermediate data, by patch id
    StepLog log;
    u64 step_count = 0; // SolverLog::step_count (SolverLog.hpp:49-54): steps registered so far

    static constexpr u64 GRID = 1ull << 21; // PatchScheduler::max_axis_patch_coord_length

    /// ref: CoordRangeTransform<u64_3,f64_3> "multiply" mode (CoordRangeTransform.cpp:95-110,
    /// CoordRangeTransform.hpp:112-120): obj = f64(pc - 0) * fact + bmin, fact = (bmax-bmin)/2^21
    void patch_box(const Patch &p, f64 lo[3], f64 hi[3]) const {
        for (int d = 0; d < 3; d++) {
            f64 fact = (box_max[d] - box_min[d]) / f64(GRID);
            lo[d]    = f64(p.coord_min[d]) * fact + box_min[d];
            hi[d]    = f64(p.coord_max[d] + 1) * fact + box_min[d];
        }
    }

    /// ref: shammodels/sph/src/modules/ParticleReordering.cpp:22-51 — per patch: Morton codes of the
    /// positions over the PATCH box (shamtree/src/RadixTreeMortonBuilder.cpp:68-107: codes, pad to a power of
    /// two with the error code, index buffer, key/value bitonic sort), then PatchDataLayer::index_remap:
    /// new[i] = old[index_map[i]] for every field
    void reorder_particles() {
        for (auto &p : patches) {
            u32 n = p.pdat.n;
            if (n == 0)
                continue;
            f64 lo[3], hi[3];
            patch_box(p, lo, hi);
            u32 P2 = roundup_pow2(n);
            std::vector<u32> codes(P2), ids(P2);
            morton_code_set_from_positions<u32>(p.pdat.xyz.data(), 3, n, lo, hi, P2, codes.data());
            for (u32 i = 0; i < P2; i++)
                ids[i] = i;
            sort_by_key_bitonic<u32>(codes.data(), ids.data(), P2);
            ids.resize(n);
            for (u32 i : ids)
                if (i >= n)
                    throw std::runtime_error("reorder_particles: a padding entry sorted below a particle");
            p.pdat.keep_ids(ids);
        }
    }

    /// static grid of nx*ny*nz patches (powers of two), id = x + nx*(y + ny*z)
    void init_patch_grid(u32 nx, u32 ny, u32 nz) {
        patches.clear();
        u32 nn[3] = {nx, ny, nz};
        for (u32 z = 0; z < nz; z++)
            for (u32 y = 0; y < ny; y++)
                for (u32 x = 0; x < nx; x++) {
                    Patch p;
                    p.id      = x + nx * (y + u64(ny) * z);
                    u32 c[3]  = {x, y, z};
                    for (int d = 0; d < 3; d++) {
                        u64 sz          = GRID / nn[d];
                        p.coord_min[d]  = sz * c[d];
                        p.coord_max[d]  = sz * (c[d] + 1) - 1;
                    }
                    patches.push_back(std::move(p));
                }
        next_patch_id = u64(nx) * ny * nz;
    }

    /// owner patch of a position (SerialPatchTree::compute_patch_owner semantics: the patch
    /// whose [lo, hi) box contains the point); returns index in `patches` or -1
    int patch_owner(const f64 r[3]) const {
        for (size_t k = 0; k < patches.size(); k++) {
            f64 lo[3], hi[3];
            patch_box(patches[k], lo, hi);
            if (lo[0] <= r[0] && r[0] < hi[0] && lo[1] <= r[1] && r[1] < hi[1] && lo[2] <= r[2]
                && r[2] < hi[2])
                return int(k);
        }
        return -1;
    }

    u64 total_count() const {
        u64 s = 0;
        for (auto &p : patches)
            s += p.pdat.n;
        return s;
    }

    u64 next_patch_id = 0; ///< SchedulerPatchList::_next_patch_id (set by init_patch_grid)

    /// PatchScheduler::split_patches for one patch.  ref: shamrock/include/shamrock/patch/PatchCoord.hpp:36-120
    /// (split coordinate = ((max - min + 1) / 2) - 1 + min per axis; child c = 4 ix + 2 iy + iz),
    /// shamrock/src/scheduler/scheduler_patch_list.cpp:109-157 (child 0 keeps the id and the place of the
    /// parent, children 1..7 get new ids and are appended), SchedulerPatchData.cpp:302-358 +
    /// PatchDataLayer::split_patchdata (every object goes to the child whose [lo, hi) box holds it, order kept)
    void split_patch(u32 ip) {
        Patch parent = std::move(patches.at(ip));
        u64 sp[3];
        for (int d = 0; d < 3; d++) {
            if (parent.coord_max[d] == parent.coord_min[d])
                throw std::runtime_error("patch cannot be split any further");
            sp[d] = ((parent.coord_max[d] - parent.coord_min[d]) + 1) / 2 - 1 + parent.coord_min[d];
        }
        Patch ch[8];
        for (int c = 0; c < 8; c++) {
            int up[3]  = {(c >> 2) & 1, (c >> 1) & 1, c & 1};
            ch[c].id   = c == 0 ? parent.id : next_patch_id++;
            for (int d = 0; d < 3; d++) {
                ch[c].coord_min[d] = up[d] ? sp[d] + 1 : parent.coord_min[d];
                ch[c].coord_max[d] = up[d] ? parent.coord_max[d] : sp[d];
            }
            ch[c].pdat.resize(0);
        }
        std::vector<std::vector<u32>> ids(8);
        for (u32 i = 0; i < parent.pdat.n; i++) {
            const f64 *r = &parent.pdat.xyz[3 * size_t(i)];
            int own      = -1;
            for (int c = 0; c < 8 && own < 0; c++) {
                f64 lo[3], hi[3];
                patch_box(ch[c], lo, hi);
                if (lo[0] <= r[0] && r[0] < hi[0] && lo[1] <= r[1] && r[1] < hi[1] && lo[2] <= r[2] && r[2] < hi[2])
                    own = c;
            }
            if (own < 0)
                throw std::runtime_error("split_patchdata: an object is outside of the patch that is split");
            ids[own].push_back(i);
        }
        for (int c = 0; c < 8; c++)
            ch[c].pdat.append_subset_from(parent.pdat, ids[c]);
        patches[ip] = std::move(ch[0]);
        for (int c = 1; c < 8; c++)
            patches.push_back(std::move(ch[c]));
        step.clear();
    }

    /// PatchScheduler::merge_patches for the octet whose child 0 is patch `ip0`.  ref:
    /// scheduler_patch_list.cpp:160-185 + Patch::merge_patch (Patch.hpp:253-286: the merged patch keeps the id
    /// of child 0), SchedulerPatchData::merge_patchdata (children appended in child order 0..7)
    void merge_patches(u32 ip0) {
        Patch &p0 = patches.at(ip0);
        u64 ext[3];
        for (int d = 0; d < 3; d++)
            ext[d] = p0.coord_max[d] - p0.coord_min[d] + 1;
        std::vector<size_t> sib(8, size_t(-1));
        for (int c = 0; c < 8; c++) {
            int up[3] = {(c >> 2) & 1, (c >> 1) & 1, c & 1};
            for (size_t k = 0; k < patches.size(); k++) {
                bool ok = true;
                for (int d = 0; d < 3; d++)
                    ok = ok && patches[k].coord_min[d] == p0.coord_min[d] + up[d] * ext[d]
                         && patches[k].coord_max[d] == p0.coord_max[d] + up[d] * ext[d];
                if (ok)
                    sib[c] = k;
            }
            if (sib[c] == size_t(-1))
                throw std::runtime_error("merge_patches: the octet of siblings is not complete");
        }
        for (int c = 1; c < 8; c++) {
            PatchData &src = patches[sib[c]].pdat;
            std::vector<u32> all(src.n);
            for (u32 i = 0; i < src.n; i++)
                all[i] = i;
            p0.pdat.append_subset_from(src, all);
        }
        for (int d = 0; d < 3; d++)
            p0.coord_max[d] = p0.coord_min[d] + 2 * ext[d] - 1;
        std::vector<size_t> dead(sib.begin() + 1, sib.end());
        std::sort(dead.begin(), dead.end());
        for (size_t q = dead.size(); q-- > 0;)
            patches.erase(patches.begin() + dead[q]);
        step.clear();
    }

    void evolve_once();
};

// -------------------------------------------------------------------------------------------
template<class K>
struct SolverT {
    using Kern = SPHKernel<K>;
    Solver &S;
    SolverConfig &cfg;
    explicit SolverT(Solver &s) : S(s), cfg(s.cfg) {}

    // ---- streaming pieces -----------------------------------------------------------------
    /// ForwardEuler node: field = field + dt * derivative
    /// ref: shammodels/common/include/shammodels/common/modules/ForwardEuler.hpp:62-69
    static void forward_euler(std::vector<f64> &f, const std::vector<f64> &d, f64 dt) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t) f.size(); i++)
            f[i] = f[i] + dt * d[i];
    }

    /// ref: shammodels/sph/src/Solver.cpp:390-524 ("leapfrog predictor" sequence:
    /// half_step1{vxyz,uint}, full_step_xyz, half_step2{vxyz,uint}), dt_half = dt/2 (:582-594)
    void predictor(f64 dt) {
        f64 dt_half = dt / 2;
        for (auto &p : S.patches) {
            auto &d = p.pdat;
            forward_euler(d.vxyz, d.axyz, dt_half);
            forward_euler(d.uint, d.duint, dt_half);
            forward_euler(d.xyz, d.vxyz, dt);
            forward_euler(d.vxyz, d.axyz, dt_half);
            forward_euler(d.uint, d.duint, dt_half);
        }
    }

    /// kill spheres.  ref: shammodels/sph/src/Solver.cpp:526-578; modules/
    /// GetParticlesOutsideSphere.cpp:25-45 (|r - c| > radius) → KillParticles.cpp:22-35 →
    /// PatchDataField::remove_ids (shamrock/src/patch/PatchDataField.cpp:300-331, order kept)
    void kill_particles() {
        for (int s = 0; s < cfg.n_kill_spheres; s++) {
            for (auto &p : S.patches) {
                auto &d = p.pdat;
                std::vector<u32> keep;
                for (u32 i = 0; i < d.n; i++) {
                    vec3 r{d.xyz[3 * i] - cfg.kill_center[s][0], d.xyz[3 * i + 1] - cfg.kill_center[s][1],
                           d.xyz[3 * i + 2] - cfg.kill_center[s][2]};
                    bool outside = length(r) > cfg.kill_radius[s];
                    if (!outside)
                        keep.push_back(i);
                }
                if (keep.size() != d.n)
                    d.keep_ids(keep);
            }
        }
    }

    /// ref: shammodels/sph/src/modules/ExternalForces.cpp:593-700 (accretion onto the central
    /// point mass: particles with |r|^2 <= Racc^2 are removed)
    void point_mass_accrete() {
        if (!cfg.has_point_mass)
            return;
        for (auto &p : S.patches) {
            auto &d = p.pdat;
            std::vector<u32> keep;
            f64 acc_rad2 = cfg.pm_racc * cfg.pm_racc;
            for (u32 i = 0; i < d.n; i++) {
                vec3 r{d.xyz[3 * i], d.xyz[3 * i + 1], d.xyz[3 * i + 2]};
                if (dot(r, r) > acc_rad2)
                    keep.push_back(i);
            }
            if (keep.size() != d.n)
                d.keep_ids(keep);
        }
    }

    /// ref: ExternalForces.cpp:49-323 (axyz_ext reset + point mass:
    /// shammodels/common/src/modules/AddForceCentralGravPotential.cpp:38-48)
    void compute_ext_forces_indep_v() {
        for (auto &p : S.patches) {
            auto &d = p.pdat;
            std::fill(d.axyz_ext.begin(), d.axyz_ext.end(), 0.);
            if (cfg.has_point_mass) {
                f64 mGM = -cfg.pm_mass * cfg.constant_G;
#pragma omp parallel for schedule(static)
                for (int64_t i = 0; i < (int64_t) d.n; i++) {
                    vec3 r_a{d.xyz[3 * i], d.xyz[3 * i + 1], d.xyz[3 * i + 2]};
                    r_a          = r_a - vec3{0, 0, 0};
                    f64 abs_ra   = length(r_a);
                    f64 abs_ra_3 = abs_ra * abs_ra * abs_ra;
                    vec3 inc     = mGM * r_a / abs_ra_3;
                    d.axyz_ext[3 * i] += inc.x;
                    d.axyz_ext[3 * i + 1] += inc.y;
                    d.axyz_ext[3 * i + 2] += inc.z;
                }
            }
        }
    }

    /// periodic wrap.  ref: shamrock/src/math/integrators.cpp:207-236
    void apply_position_boundary() {
        if (cfg.bc == BC_PERIODIC) {
            for (auto &p : S.patches) {
                auto &d = p.pdat;
#pragma omp parallel for schedule(static)
                for (int64_t i = 0; i < (int64_t) d.n; i++)
                    for (int c = 0; c < 3; c++) {
                        f64 delt = S.box_max[c] - S.box_min[c];
                        f64 r    = d.xyz[3 * i + c] - S.box_min[c];
                        r        = std::fmod(r, delt);
                        r += delt;
                        r = std::fmod(r, delt);
                        r += S.box_min[c];
                        d.xyz[3 * i + c] = r;
                    }
            }
        }
        reattribute();
    }

    /// ref: shamrock/include/shamrock/scheduler/ReattributeDataUtility.hpp:40-230 — kept ids stay
    /// in order, migrants are appended to their new patch grouped by (sender id asc), in index order
    void reattribute() {
        if (S.patches.size() == 1 && cfg.bc == BC_PERIODIC)
            return; // everything is inside the single patch after the wrap
        size_t np = S.patches.size();
        std::vector<std::map<size_t, std::vector<u32>>> extract(np);
        std::vector<std::vector<u32>> keep(np);
        bool any = false;
        for (size_t k = 0; k < np; k++) {
            auto &d = S.patches[k].pdat;
            for (u32 i = 0; i < d.n; i++) {
                int own = S.patch_owner(&d.xyz[3 * i]);
                if (own < 0) {
                    if (np == 1)
                        own = 0; // free boundaries, single patch: the box is not a constraint
                    else
                        throw std::runtime_error("a new id could not be computed");
                }
                if (size_t(own) != k) {
                    extract[k][size_t(own)].push_back(i);
                    any = true;
                } else
                    keep[k].push_back(i);
            }
        }
        if (!any)
            return;
        std::vector<PatchData> old(np);
        for (size_t k = 0; k < np; k++)
            old[k] = S.patches[k].pdat;
        for (size_t k = 0; k < np; k++)
            if (keep[k].size() != old[k].n)
                S.patches[k].pdat.keep_ids(keep[k]);
        for (size_t k = 0; k < np; k++)       // sender, ascending id
            for (auto &[dst, ids] : extract[k]) // receiver
                S.patches[dst].pdat.append_subset_from(old[k], ids);
    }

    // ---- ghosts ----------------------------------------------------------------------------
    /// ref: shammodels/sph/include/shammodels/sph/SPHUtilities.hpp:74-103 (interactR_patch =
    /// max(h)*htol*Rkern), shammodels/sph/src/BasicSPHGhosts.cpp:261-509 (find_interfaces),
    /// :512-579 (gen_id_table_interfaces; ids = stream compaction → ascending)
    std::vector<Interface> build_ghost_cache() {
        size_t np = S.patches.size();
        std::vector<f64> interactR(np);
        for (size_t k = 0; k < np; k++) {
            auto &d = S.patches[k].pdat;
            if (d.n > 0) {
                f64 hm = d.hpart[0];
                for (u32 i = 1; i < d.n; i++)
                    hm = std::fmax(hm, d.hpart[i]);
                interactR[k] = hm * cfg.htol_up_coarse_cycle * Kern::Rkern;
            } else
                interactR[k] = std::numeric_limits<f64>::lowest();
        }
        f64 bsize[3] = {S.box_max[0] - S.box_min[0], S.box_max[1] - S.box_min[1],
                        S.box_max[2] - S.box_min[2]};
        int rep      = (cfg.bc == BC_PERIODIC) ? 1 : 0;
        // multimap<(sender,receiver)> : key order, equal keys in insertion (offset-loop) order
        std::map<std::pair<u64, u64>, std::vector<Interface>> mm;
        for (i32 xoff = -rep; xoff <= rep; xoff++)
            for (i32 yoff = -rep; yoff <= rep; yoff++)
                for (i32 zoff = -rep; zoff <= rep; zoff++) {
                    f64 off[3] = {xoff * bsize[0], yoff * bsize[1], zoff * bsize[2]};
                    for (size_t s = 0; s < np; s++) {
                        if (S.patches[s].pdat.n == 0)
                            continue; // (empty sender: get_ids_where gives 0 → skipped anyway)
                        f64 slo[3], shi[3];
                        S.patch_box(S.patches[s], slo, shi);
                        for (size_t r = 0; r < np; r++) {
                            if (S.patches[r].pdat.n == 0)
                                continue;
                            if (r == s && xoff == 0 && yoff == 0 && zoff == 0)
                                continue;
                            f64 rlo[3], rhi[3];
                            S.patch_box(S.patches[r], rlo, rhi);
                            f64 R  = interactR[r];
                            bool ok = true;
                            Interface itf;
                            for (int d = 0; d < 3; d++) {
                                f64 elo = rlo[d] - R, ehi = rhi[d] + R; // receiv_exp
                                f64 so_lo = slo[d] + off[d], so_hi = shi[d] + off[d];
                                f64 ilo = std::fmax(elo, so_lo), ihi = std::fmin(ehi, so_hi);
                                if (!(ihi >= ilo))
                                    ok = false;
                                // interf_volume = sender ∩ (receiv_exp + (-off))
                                f64 moff     = -off[d];
                                itf.cut_lo[d] = std::fmax(slo[d], elo + moff);
                                itf.cut_hi[d] = std::fmin(shi[d], ehi + moff);
                                itf.offset[d] = off[d];
                            }
                            if (!ok)
                                continue;
                            itf.sender   = S.patches[s].id;
                            itf.receiver = S.patches[r].id;
                            itf.ioff[0]  = xoff;
                            itf.ioff[1]  = yoff;
                            itf.ioff[2]  = zoff;
                            auto &d      = S.patches[s].pdat;
                            for (u32 i = 0; i < d.n; i++) {
                                const f64 *x = &d.xyz[3 * i];
                                if (itf.cut_lo[0] <= x[0] && x[0] < itf.cut_hi[0]
                                    && itf.cut_lo[1] <= x[1] && x[1] < itf.cut_hi[1]
                                    && itf.cut_lo[2] <= x[2] && x[2] < itf.cut_hi[2])
                                    itf.ids.push_back(i);
                            }
                            if (itf.ids.empty())
                                continue;
                            mm[{itf.sender, itf.receiver}].push_back(std::move(itf));
                        }
                    }
                }
        std::vector<Interface> out;
        for (auto &[k, v] : mm)
            for (auto &i : v)
                out.push_back(std::move(i));
        return out;
    }

    Patch &patch_by_id(u64 id) {
        for (auto &p : S.patches)
            if (p.id == id)
                return p;
        throw std::runtime_error("unknown patch id");
    }

    /// ref: shammodels/sph/include/shammodels/sph/BasicSPHGhosts.hpp:294-321,476-514
    void merge_position_ghost(const std::vector<Interface> &itfs) {
        for (auto &p : S.patches) {
            if (p.pdat.n == 0)
                continue;
            PatchStep &st = S.step[p.id];
            st.n          = p.pdat.n;
            st.mxyz       = p.pdat.xyz;
            st.mh         = p.pdat.hpart;
        }
        for (auto &itf : itfs) { // key order (sender, receiver): for a given receiver → sender asc
            PatchStep &st = S.step[itf.receiver];
            auto &sd      = patch_by_id(itf.sender).pdat;
            for (u32 id : itf.ids) {
                for (int c = 0; c < 3; c++)
                    st.mxyz.push_back(sd.xyz[3 * id + c] + itf.offset[c]);
                st.mh.push_back(sd.hpart[id]);
            }
        }
        for (auto &[id, st] : S.step)
            st.m = u32(st.mh.size());
    }

    /// ref: shammodels/sph/src/modules/BuildTrees.cpp:26-66
    void build_merged_pos_trees() {
        for (auto &[id, st] : S.step) {
            f64 bmin[3], bmax[3];
            for (int c = 0; c < 3; c++) {
                f64 mn = st.mxyz[c], mx = st.mxyz[c];
                for (u32 i = 1; i < st.m; i++) {
                    mn = std::fmin(mn, st.mxyz[3 * i + c]);
                    mx = std::fmax(mx, st.mxyz[3 * i + c]);
                }
                const f64 inf = std::numeric_limits<f64>::infinity();
                bmin[c]       = std::nextafter(mn, -inf);
                bmax[c]       = std::nextafter(mx, inf);
            }
            st.tree = rebuild_from_positions<u32>(
                st.mxyz.data(), 3, st.m, bmin, bmax, cfg.tree_reduction_level);
        }
    }

    /// ref: shammodels/sph/src/Solver.cpp:1322-1356
    void compute_presteps_rint() {
        for (auto &[id, st] : S.step) {
            st.rint = compute_tree_field_max_field(st.tree, st.mh.data());
            for (auto &v : st.rint)
                v *= cfg.htol_up_coarse_cycle;
        }
    }

    void start_neighbors_cache() {
        for (auto &[id, st] : S.step) {
            xyzh_view P{st.mxyz.data(), 3, st.mh.data()};
            if (cfg.use_two_stage_search)
                st.cache = neighbour_cache_2stages(
                    st.tree, P, st.n, st.rint, Kern::Rkern, cfg.htol_up_coarse_cycle);
            else
                st.cache = neighbour_cache_1stage(
                    st.tree, P, st.n, st.rint, Kern::Rkern, cfg.htol_up_coarse_cycle);
        }
    }

    /// ref: shammodels/sph/src/Solver.cpp:1060-1304, modules/LoopSmoothingLengthIter.cpp:29-84
    void sph_prestep() {
        u32 hstep_cnt = 0;
        for (; hstep_cnt < cfg.h_max_subcycles_count; hstep_cnt++) {
            S.step.clear();
            auto itfs = build_ghost_cache();
            merge_position_ghost(itfs);
            build_merged_pos_trees();
            compute_presteps_rint();
            start_neighbors_cache();
            interfaces = std::move(itfs);

            if (cfg.gpart_mass == 0)
                throw std::runtime_error("invalid gpart_mass 0, this configuration can not converge");

            std::map<u64, std::vector<f64>> eps_h, h_old;
            for (auto &p : S.patches) {
                if (p.pdat.n == 0)
                    continue;
                eps_h[p.id].assign(p.pdat.n, 100.);
                h_old[p.id] = p.pdat.hpart;
            }
            f64 local_max_eps = std::numeric_limits<f64>::max();
            u32 iter_h        = 0;
            for (; iter_h < cfg.h_iter_per_subcycles; iter_h++) {
                for (auto &p : S.patches) {
                    if (p.pdat.n == 0)
                        continue;
                    PatchStep &st = S.step[p.id];
                    iterate_smoothing_length_density<K>(
                        st.cache, st.mxyz.data(), 3, st.n, h_old[p.id].data(), p.pdat.hpart.data(),
                        eps_h[p.id].data(), cfg.gpart_mass, cfg.htol_up_coarse_cycle,
                        cfg.htol_up_fine_cycle);
                }
                local_max_eps = std::numeric_limits<f64>::lowest();
                for (auto &[id, e] : eps_h)
                    for (f64 v : e)
                        local_max_eps = std::fmax(local_max_eps, v);
                if (local_max_eps < cfg.epsilon_h)
                    break;
            }
            S.log.h_iters_last = iter_h;
            f64 local_min_eps  = std::numeric_limits<f64>::max();
            for (auto &[id, e] : eps_h)
                for (f64 v : e)
                    local_min_eps = std::fmin(local_min_eps, v);
            bool should_rerun_gz = local_min_eps < 0;
            bool below_tol       = local_max_eps < cfg.epsilon_h;
            bool converged       = below_tol && !should_rerun_gz;
            if (!converged)
                continue;
            break;
        }
        S.log.h_subcycles = hstep_cnt + 1;
        for (auto &p : S.patches) {
            if (p.pdat.n == 0)
                continue;
            PatchStep &st = S.step[p.id];
            st.omega.resize(st.n);
            compute_omega<K>(
                st.cache, st.mxyz.data(), 3, st.n, p.pdat.hpart.data(), st.omega.data(),
                cfg.gpart_mass);
        }
    }

    std::vector<Interface> interfaces;

    // ---- corrector-loop pieces --------------------------------------------------------------
    /// ref: shammodels/sph/src/Solver.cpp:1394-1633 (ghost layout: SolverConfig.cpp:124-165)
    void communicate_merge_ghosts_fields() {
        for (auto &p : S.patches) {
            if (p.pdat.n == 0)
                continue;
            PatchStep &st = S.step[p.id];
            st.g_h        = p.pdat.hpart;
            st.g_u        = p.pdat.uint;
            st.g_v        = p.pdat.vxyz;
            st.g_omega    = st.omega;
            if (cfg.has_axyz_in_ghost())
                st.g_a = p.pdat.axyz;
            if (cfg.eos == EOS_LOCALLY_ISOTHERMAL_LP07)
                st.g_cs = p.pdat.soundspeed;
        }
        for (auto &itf : interfaces) {
            PatchStep &st  = S.step[itf.receiver];
            auto &sp       = patch_by_id(itf.sender);
            auto &sd       = sp.pdat;
            PatchStep &sst = S.step[itf.sender];
            for (u32 id : itf.ids) {
                st.g_h.push_back(sd.hpart[id]);
                st.g_u.push_back(sd.uint[id]);
                for (int c = 0; c < 3; c++)
                    st.g_v.push_back(sd.vxyz[3 * id + c]);
                st.g_omega.push_back(sst.omega[id]);
                if (cfg.has_axyz_in_ghost())
                    for (int c = 0; c < 3; c++)
                        st.g_a.push_back(sd.axyz[3 * id + c]);
                if (cfg.eos == EOS_LOCALLY_ISOTHERMAL_LP07)
                    st.g_cs.push_back(sd.soundspeed[id]);
            }
        }
    }

    /// ref: shammodels/sph/src/modules/DiffOperator.cpp:25-146 (divv), :148-270 (curlv)
    void update_divv_curlv(bool do_curl) {
        for (auto &p : S.patches) {
            if (p.pdat.n == 0)
                continue;
            PatchStep &st    = S.step[p.id];
            const f64 pmass  = cfg.gpart_mass;
            const f64 Rker2  = Kern::Rkern * Kern::Rkern;
            auto &c          = st.cache;
#pragma omp parallel for schedule(dynamic, 256)
            for (int64_t ia = 0; ia < (int64_t) st.n; ia++) {
                u32 id_a = u32(ia);
                f64 h_a  = st.g_h[id_a];
                vec3 xyz_a{st.mxyz[3 * id_a], st.mxyz[3 * id_a + 1], st.mxyz[3 * id_a + 2]};
                vec3 vxyz_a{st.g_v[3 * id_a], st.g_v[3 * id_a + 1], st.g_v[3 * id_a + 2]};
                f64 omega_a         = st.g_omega[id_a];
                f64 rho_a           = rho_h(pmass, h_a, Kern::hfactd);
                f64 inv_rho_omega_a = 1. / (omega_a * rho_a);
                f64 sum_nabla_v     = 0;
                vec3 sum_nabla_cross_v{0, 0, 0};
                u32 s0 = c.scanned_cnt[id_a], s1 = s0 + c.cnt_neigh[id_a];
                for (u32 k = s0; k < s1; k++) {
                    u32 id_b = c.index_neigh_map[k];
                    vec3 dr  = xyz_a - vec3{st.mxyz[3 * id_b], st.mxyz[3 * id_b + 1], st.mxyz[3 * id_b + 2]};
                    f64 rab2 = dot(dr, dr);
                    f64 h_b  = st.g_h[id_b];
                    if (rab2 > h_a * h_a * Rker2 && rab2 > h_b * h_b * Rker2)
                        continue;
                    f64 rab = std::sqrt(rab2);
                    vec3 vxyz_b{st.g_v[3 * id_b], st.g_v[3 * id_b + 1], st.g_v[3 * id_b + 2]};
                    vec3 v_ab      = vxyz_a - vxyz_b;
                    vec3 r_ab_unit = dr / rab;
                    if (rab < 1e-9)
                        r_ab_unit = {0, 0, 0};
                    vec3 dWab_a = Kern::dW_3d(rab, h_a) * r_ab_unit;
                    sum_nabla_v += pmass * dot(v_ab, dWab_a);
                    if (do_curl)
                        sum_nabla_cross_v += pmass * cross(v_ab, dWab_a);
                }
                p.pdat.divv[id_a] = -inv_rho_omega_a * sum_nabla_v;
                if (do_curl) {
                    vec3 cv                     = -inv_rho_omega_a * sum_nabla_cross_v;
                    p.pdat.curlv[3 * id_a]     = cv.x;
                    p.pdat.curlv[3 * id_a + 1] = cv.y;
                    p.pdat.curlv[3 * id_a + 2] = cv.z;
                }
            }
        }
    }

    // ref: shammath/include/shammath/matrix_legacy.hpp:24-92
    static std::array<vec3, 3> compute_inv_33(std::array<vec3, 3> mat) {
        f64 a00 = mat[0].x, a10 = mat[1].x, a20 = mat[2].x;
        f64 a01 = mat[0].y, a11 = mat[1].y, a21 = mat[2].y;
        f64 a02 = mat[0].z, a12 = mat[1].z, a22 = mat[2].z;
        f64 det
            = (-a02 * a11 * a20 + a01 * a12 * a20 + a02 * a10 * a21 - a00 * a12 * a21
               - a01 * a10 * a22 + a00 * a11 * a22);
        return {
            (vec3{-a12 * a21 + a11 * a22, a02 * a21 - a01 * a22, -a02 * a11 + a01 * a12} / det),
            (vec3{a12 * a20 - a10 * a22, -a02 * a20 + a00 * a22, a02 * a10 - a00 * a12} / det),
            (vec3{-a11 * a20 + a10 * a21, a01 * a20 - a00 * a21, -a01 * a10 + a00 * a11} / det)};
    }
    static std::array<vec3, 3> mat_prod_33(std::array<vec3, 3> A, std::array<vec3, 3> B) {
        f64 a00 = A[0].x, a10 = A[1].x, a20 = A[2].x;
        f64 a01 = A[0].y, a11 = A[1].y, a21 = A[2].y;
        f64 a02 = A[0].z, a12 = A[1].z, a22 = A[2].z;
        f64 b00 = B[0].x, b10 = B[1].x, b20 = B[2].x;
        f64 b01 = B[0].y, b11 = B[1].y, b21 = B[2].y;
        f64 b02 = B[0].z, b12 = B[1].z, b22 = B[2].z;
        return {
            vec3{a00 * b00 + a01 * b10 + a02 * b20, a00 * b01 + a01 * b11 + a02 * b21,
                 a00 * b02 + a01 * b12 + a02 * b22},
            vec3{a10 * b00 + a11 * b10 + a12 * b20, a10 * b01 + a11 * b11 + a12 * b21,
                 a10 * b02 + a11 * b12 + a12 * b22},
            vec3{a20 * b00 + a21 * b10 + a22 * b20, a20 * b01 + a21 * b11 + a22 * b21,
                 a20 * b02 + a21 * b12 + a22 * b22}};
    }

    /// ref: shammodels/sph/src/modules/DiffOperatorDtDivv.cpp:100-196 (also_do_div_curl_v=false)
    /// and :220-353 (true: also writes divv / curlv from the matrix form)
    void update_dtdivv(bool also_do_div_curl_v) {
        for (auto &p : S.patches) {
            if (p.pdat.n == 0)
                continue;
            PatchStep &st   = S.step[p.id];
            const f64 pmass = cfg.gpart_mass;
            const f64 Rker2 = Kern::Rkern * Kern::Rkern;
            auto &c         = st.cache;
#pragma omp parallel for schedule(dynamic, 256)
            for (int64_t ia = 0; ia < (int64_t) st.n; ia++) {
                u32 id_a = u32(ia);
                f64 h_a  = st.g_h[id_a];
                vec3 xyz_a{st.mxyz[3 * id_a], st.mxyz[3 * id_a + 1], st.mxyz[3 * id_a + 2]};
                vec3 vxyz_a{st.g_v[3 * id_a], st.g_v[3 * id_a + 1], st.g_v[3 * id_a + 2]};
                vec3 axyz_a{st.g_a[3 * id_a], st.g_a[3 * id_a + 1], st.g_a[3 * id_a + 2]};
                const vec3 Z{0, 0, 0};
                std::array<vec3, 3> Rij_a{Z, Z, Z}, Rij_a_dvk_dxj{Z, Z, Z}, Rij_a_dak_dxj{Z, Z, Z};
                u32 s0 = c.scanned_cnt[id_a], s1 = s0 + c.cnt_neigh[id_a];
                for (u32 k = s0; k < s1; k++) {
                    u32 id_b  = c.index_neigh_map[k];
                    vec3 r_ab = xyz_a - vec3{st.mxyz[3 * id_b], st.mxyz[3 * id_b + 1], st.mxyz[3 * id_b + 2]};
                    f64 rab2  = dot(r_ab, r_ab);
                    f64 h_b   = st.g_h[id_b];
                    if (rab2 > h_a * h_a * Rker2 && rab2 > h_b * h_b * Rker2)
                        continue;
                    f64 rab = std::sqrt(rab2);
                    vec3 vxyz_b{st.g_v[3 * id_b], st.g_v[3 * id_b + 1], st.g_v[3 * id_b + 2]};
                    vec3 axyz_b{st.g_a[3 * id_b], st.g_a[3 * id_b + 1], st.g_a[3 * id_b + 2]};
                    vec3 v_ab      = vxyz_a - vxyz_b;
                    vec3 a_ab      = axyz_a - axyz_b;
                    vec3 r_ab_unit = r_ab / rab;
                    if (rab < 1e-9)
                        r_ab_unit = {0, 0, 0};
                    vec3 dWab_a  = Kern::dW_3d(rab, h_a) * r_ab_unit;
                    vec3 mdWab_b = dWab_a * pmass;
                    Rij_a[0] -= r_ab.x * mdWab_b;
                    Rij_a[1] -= r_ab.y * mdWab_b;
                    Rij_a[2] -= r_ab.z * mdWab_b;
                    Rij_a_dvk_dxj[0] -= v_ab * mdWab_b.x;
                    Rij_a_dvk_dxj[1] -= v_ab * mdWab_b.y;
                    Rij_a_dvk_dxj[2] -= v_ab * mdWab_b.z;
                    Rij_a_dak_dxj[0] -= a_ab * mdWab_b.x;
                    Rij_a_dak_dxj[1] -= a_ab * mdWab_b.y;
                    Rij_a_dak_dxj[2] -= a_ab * mdWab_b.z;
                }
                auto invRij  = compute_inv_33(Rij_a);
                auto dvi_dxk = mat_prod_33(invRij, Rij_a_dvk_dxj);
                auto dai_dxk = mat_prod_33(invRij, Rij_a_dak_dxj);
                f64 div_ai   = dai_dxk[0].x + dai_dxk[1].y + dai_dxk[2].z;
                f64 div_vi   = dvi_dxk[0].x + dvi_dxk[1].y + dvi_dxk[2].z;
                vec3 curl_vi = {dvi_dxk[1].z - dvi_dxk[2].y, dvi_dxk[2].x - dvi_dxk[0].z,
                                dvi_dxk[0].y - dvi_dxk[1].x};
                f64 tens_nablav
                    = dvi_dxk[0].x * dvi_dxk[0].x + dvi_dxk[1].x * dvi_dxk[0].y
                      + dvi_dxk[2].x * dvi_dxk[0].z + dvi_dxk[0].y * dvi_dxk[1].x
                      + dvi_dxk[1].y * dvi_dxk[1].y + dvi_dxk[2].y * dvi_dxk[1].z
                      + dvi_dxk[0].z * dvi_dxk[2].x + dvi_dxk[1].z * dvi_dxk[2].y
                      + dvi_dxk[2].z * dvi_dxk[2].z;
                if (also_do_div_curl_v) {
                    p.pdat.divv[id_a]          = div_vi;
                    p.pdat.curlv[3 * id_a]     = curl_vi.x;
                    p.pdat.curlv[3 * id_a + 1] = curl_vi.y;
                    p.pdat.curlv[3 * id_a + 2] = curl_vi.z;
                }
                p.pdat.dtdivv[id_a] = div_ai - tens_nablav;
            }
        }
    }

    /// ref: shammodels/sph/src/modules/UpdateViscosity.cpp:52-119 (MM97), :122-222 (CD10)
    void update_artificial_viscosity(f64 dt) {
        if (!cfg.has_alphaAV())
            return;
        const f64 eps_d = std::numeric_limits<f64>::epsilon();
        for (auto &p : S.patches) {
            if (p.pdat.n == 0)
                continue;
            PatchStep &st = S.step[p.id];
            auto &d       = p.pdat;
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < (int64_t) d.n; i++) {
                f64 cs_a = d.soundspeed[i], h_a = d.hpart[i], alpha_a = d.alpha_AV[i];
                f64 divv_a          = d.divv[i];
                f64 vsig            = cs_a;
                f64 inv_tau_a       = vsig * cfg.sigma_decay / h_a;
                f64 fact_t          = dt * inv_tau_a;
                f64 euler_impl_fact = 1 / (1 + fact_t);
                if (cfg.av == AV_MM97) {
                    f64 source    = std::fmax(0., -divv_a);
                    f64 new_alpha = (alpha_a + source * dt + fact_t * cfg.alpha_min) * euler_impl_fact;
                    st.alpha_updated[i] = std::fmin(cfg.alpha_max, new_alpha);
                } else {
                    vec3 curlv_a{d.curlv[3 * i], d.curlv[3 * i + 1], d.curlv[3 * i + 2]};
                    f64 dtdivv_a = d.dtdivv[i];
                    f64 fac      = std::fmax(-divv_a, 0.);
                    fac *= fac;
                    f64 traceS        = dot(curlv_a, curlv_a);
                    f64 balsara_corec = (fac + traceS > eps_d) ? fac / (fac + traceS) : 1.;
                    f64 A_a           = balsara_corec * std::fmax(-dtdivv_a, 0.);
                    f64 temp          = cs_a * cs_a;
                    f64 alpha_loc_a   = std::fmin(
                        (cs_a > 0) ? 10 * h_a * h_a * A_a / (temp) : cfg.alpha_min, cfg.alpha_max);
                    alpha_loc_a   = (temp > eps_d) ? alpha_loc_a : cfg.alpha_min;
                    f64 new_alpha = (alpha_a + alpha_loc_a * fact_t) * euler_impl_fact;
                    if (alpha_loc_a > alpha_a)
                        new_alpha = alpha_loc_a;
                    st.alpha_updated[i] = new_alpha;
                }
            }
        }
    }

    /// ref: shammodels/sph/src/Solver.cpp:2325-2368
    void exchange_alpha_ghosts() {
        if (!cfg.has_alphaAV())
            return;
        for (auto &p : S.patches)
            if (p.pdat.n)
                S.step[p.id].g_alpha = S.step[p.id].alpha_updated;
        for (auto &itf : interfaces) {
            PatchStep &st  = S.step[itf.receiver];
            PatchStep &sst = S.step[itf.sender];
            for (u32 id : itf.ids)
                st.g_alpha.push_back(sst.alpha_updated[id]);
        }
    }

    /// ref: shammodels/sph/src/modules/ComputeEos.cpp:147-248 (adiabatic), :54-130 (isothermal),
    /// :724-800 (LP07), dispatcher :1138-1308; shamphys/include/shamphys/eos.hpp:20-95
    void compute_eos_fields() {
        for (auto &p : S.patches) {
            if (p.pdat.n == 0)
                continue;
            PatchStep &st = S.step[p.id];
            st.pressure.resize(st.m);
            st.soundspeed.resize(st.m);
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < (int64_t) st.m; i++) {
                f64 rho = rho_h(cfg.gpart_mass, st.g_h[i], Kern::hfactd);
                if (cfg.eos == EOS_ADIABATIC) {
                    f64 P_a          = (cfg.gamma - 1) * rho * st.g_u[i];
                    st.pressure[i]   = P_a;
                    st.soundspeed[i] = std::sqrt(cfg.gamma * P_a / rho);
                } else if (cfg.eos == EOS_ISOTHERMAL) {
                    st.pressure[i]   = cfg.cs0 * cfg.cs0 * rho;
                    st.soundspeed[i] = cfg.cs0;
                } else {
                    f64 r0sq = cfg.eos_r0 * cfg.eos_r0;
                    f64 mq   = -cfg.eos_q;
                    vec3 R{st.mxyz[3 * i], st.mxyz[3 * i + 1], st.mxyz[3 * i + 2]};
                    f64 Rsq          = dot(R, R);
                    f64 cs_sq        = (cfg.cs0 * cfg.cs0) * std::pow(Rsq / r0sq, mq);
                    st.soundspeed[i] = std::sqrt(cs_sq);
                    st.pressure[i]   = cs_sq * rho;
                }
            }
        }
    }

    // ref: shammodels/sph/include/shammodels/sph/math/forces.hpp:27-226, math/q_ab.hpp:37-62
    static f64 vsig_u_f(f64 P_a, f64 P_b, f64 rho_a, f64 rho_b) {
        f64 rho_avg = (rho_a + rho_b) * 0.5;
        f64 abs_dp  = std::fabs(P_a - P_b);
        return std::sqrt(abs_dp / rho_avg);
    }
    static f64 q_av(f64 rho, f64 vsig, f64 v_scal_rhat) {
        return std::fmax(-0.5 * rho * vsig * v_scal_rhat, 0.);
    }
    static f64 q_av_disc(f64 rho, f64 h, f64 rab, f64 alpha_av, f64 cs, f64 vsig, f64 v_scal_rhat) {
        f64 rabinv    = inv_sat_positive(rab);
        f64 prefact   = -0.5 * rho * std::fabs(rabinv) * h;
        f64 vsig_disc = (v_scal_rhat < 0.) ? vsig : (alpha_av * cs);
        return prefact * vsig_disc * v_scal_rhat;
    }
    static vec3 sph_pressure_symetric(
        f64 m_b, f64 rho_a_sq, f64 rho_b_sq, f64 P_a, f64 P_b, f64 omega_a, f64 omega_b,
        vec3 nabla_Wab_ha, vec3 nabla_Wab_hb) {
        f64 sub_fact_a = rho_a_sq * omega_a;
        f64 sub_fact_b = rho_b_sq * omega_b;
        vec3 acc_a     = ((P_a) *inv_sat_zero(sub_fact_a)) * nabla_Wab_ha;
        vec3 acc_b     = ((P_b) *inv_sat_zero(sub_fact_b)) * nabla_Wab_hb;
        return -m_b * (acc_a + acc_b);
    }
    static void add_to_derivs_sph_artif_visco_cond(
        f64 pmass, f64 rho_a_sq, f64 omega_a_rho_a_inv, f64 rho_a_inv, f64 rho_b, f64 omega_a,
        f64 omega_b, f64 Fab_a, f64 Fab_b, f64 u_a, f64 u_b, f64 P_a, f64 P_b, f64 alpha_u,
        vec3 v_ab, vec3 r_ab_unit, f64 vsig_u, f64 qa_ab, f64 qb_ab, vec3 &dv_dt, f64 &du_dt) {
        f64 AV_P_a = P_a + qa_ab;
        f64 AV_P_b = P_b + qb_ab;
        dv_dt += sph_pressure_symetric(
            pmass, rho_a_sq, rho_b * rho_b, AV_P_a, AV_P_b, omega_a, omega_b, r_ab_unit * Fab_a,
            r_ab_unit * Fab_b);
        // duint_dt_pressure
        du_dt += AV_P_a * (omega_a_rho_a_inv * rho_a_inv) * pmass * dot(v_ab, r_ab_unit * Fab_a);
        // lambda_shock_conductivity
        du_dt += pmass * alpha_u * vsig_u * (u_a - u_b) * 0.5
                 * (Fab_a * omega_a_rho_a_inv + Fab_b / (rho_b * omega_b));
    }

    /// ref: shammodels/sph/src/modules/UpdateDerivs.cpp:89-288 (constant), :580-780 (disc),
    /// modules/NodeUpdateDerivsVaryingAlphaAV.cpp:26-137 (MM97 / CD10)
    void update_derivs() {
        for (auto &p : S.patches) {
            if (p.pdat.n == 0)
                continue;
            PatchStep &st   = S.step[p.id];
            const f64 pmass = cfg.gpart_mass;
            const f64 Rker2 = Kern::Rkern * Kern::Rkern;
            auto &c         = st.cache;
            const bool varying = cfg.has_alphaAV();
            const bool disc    = cfg.av == AV_CONSTANT_DISC;
#pragma omp parallel for schedule(dynamic, 256)
            for (int64_t ia = 0; ia < (int64_t) st.n; ia++) {
                u32 id_a = u32(ia);
                f64 h_a  = st.g_h[id_a];
                vec3 xyz_a{st.mxyz[3 * id_a], st.mxyz[3 * id_a + 1], st.mxyz[3 * id_a + 2]};
                vec3 vxyz_a{st.g_v[3 * id_a], st.g_v[3 * id_a + 1], st.g_v[3 * id_a + 2]};
                f64 P_a     = st.pressure[id_a];
                f64 cs_a    = st.soundspeed[id_a];
                f64 omega_a = st.g_omega[id_a];
                f64 u_a     = st.g_u[id_a];
                f64 alpha_a = varying ? st.g_alpha[id_a] : cfg.alpha_AV;
                f64 rho_a             = rho_h(pmass, h_a, Kern::hfactd);
                f64 rho_a_sq          = rho_a * rho_a;
                f64 rho_a_inv         = 1. / rho_a;
                f64 omega_a_rho_a_inv = 1 / (omega_a * rho_a);
                vec3 force_pressure{0, 0, 0};
                f64 tmpdU_pressure = 0;
                u32 s0 = c.scanned_cnt[id_a], s1 = s0 + c.cnt_neigh[id_a];
                for (u32 k = s0; k < s1; k++) {
                    u32 id_b = c.index_neigh_map[k];
                    vec3 dr  = xyz_a - vec3{st.mxyz[3 * id_b], st.mxyz[3 * id_b + 1], st.mxyz[3 * id_b + 2]};
                    f64 rab2 = dot(dr, dr);
                    f64 h_b  = st.g_h[id_b];
                    if (rab2 > h_a * h_a * Rker2 && rab2 > h_b * h_b * Rker2)
                        continue;
                    f64 rab = std::sqrt(rab2);
                    vec3 vxyz_b{st.g_v[3 * id_b], st.g_v[3 * id_b + 1], st.g_v[3 * id_b + 2]};
                    f64 u_b     = st.g_u[id_b];
                    f64 rho_b   = rho_h(pmass, h_b, Kern::hfactd);
                    f64 P_b     = st.pressure[id_b];
                    f64 omega_b = st.g_omega[id_b];
                    f64 cs_b    = st.soundspeed[id_b];
                    f64 alpha_b = varying ? st.g_alpha[id_b] : cfg.alpha_AV;
                    f64 Fab_a   = Kern::dW_3d(rab, h_a);
                    f64 Fab_b   = Kern::dW_3d(rab, h_b);
                    vec3 v_ab          = vxyz_a - vxyz_b;
                    vec3 r_ab_unit     = dr * inv_sat_positive(rab);
                    f64 v_ab_r_ab      = dot(v_ab, r_ab_unit);
                    f64 abs_v_ab_r_ab  = std::fabs(v_ab_r_ab);
                    f64 vsig_a         = alpha_a * cs_a + cfg.beta_AV * abs_v_ab_r_ab;
                    f64 vsig_b         = alpha_b * cs_b + cfg.beta_AV * abs_v_ab_r_ab;
                    f64 vsig_u         = vsig_u_f(P_a, P_b, rho_a, rho_b);
                    f64 qa_ab, qb_ab;
                    if (disc) {
                        qa_ab = q_av_disc(rho_a, h_a, rab, alpha_a, cs_a, vsig_a, v_ab_r_ab);
                        qb_ab = q_av_disc(rho_b, h_b, rab, alpha_b, cs_b, vsig_b, v_ab_r_ab);
                    } else {
                        qa_ab = q_av(rho_a, vsig_a, v_ab_r_ab);
                        qb_ab = q_av(rho_b, vsig_b, v_ab_r_ab);
                    }
                    add_to_derivs_sph_artif_visco_cond(
                        pmass, rho_a_sq, omega_a_rho_a_inv, rho_a_inv, rho_b, omega_a, omega_b,
                        Fab_a, Fab_b, u_a, u_b, P_a, P_b, cfg.alpha_u, v_ab, r_ab_unit, vsig_u,
                        qa_ab, qb_ab, force_pressure, tmpdU_pressure);
                }
                p.pdat.axyz[3 * id_a]     = force_pressure.x;
                p.pdat.axyz[3 * id_a + 1] = force_pressure.y;
                p.pdat.axyz[3 * id_a + 2] = force_pressure.z;
                p.pdat.duint[id_a]        = tmpdU_pressure;
            }
            // add_ext_forces (ExternalForces.cpp:325-356)
            for (size_t i = 0; i < p.pdat.axyz.size(); i++)
                p.pdat.axyz[i] += p.pdat.axyz_ext[i];
        }
    }

    /// ref: shammodels/sph/src/Solver.cpp:2677-2790 (alpha_AV = 1, beta_AV = 2 hard-coded)
    void compute_vsig() {
        for (auto &p : S.patches) {
            if (p.pdat.n == 0)
                continue;
            PatchStep &st   = S.step[p.id];
            const f64 Rker2 = Kern::Rkern * Kern::Rkern;
            auto &c         = st.cache;
            st.vsig.resize(st.n);
#pragma omp parallel for schedule(dynamic, 256)
            for (int64_t ia = 0; ia < (int64_t) st.n; ia++) {
                u32 id_a = u32(ia);
                f64 h_a  = st.g_h[id_a];
                vec3 xyz_a{st.mxyz[3 * id_a], st.mxyz[3 * id_a + 1], st.mxyz[3 * id_a + 2]};
                vec3 vxyz_a{st.g_v[3 * id_a], st.g_v[3 * id_a + 1], st.g_v[3 * id_a + 2]};
                f64 cs_a     = st.soundspeed[id_a];
                f64 vsig_max = 0;
                u32 s0 = c.scanned_cnt[id_a], s1 = s0 + c.cnt_neigh[id_a];
                for (u32 k = s0; k < s1; k++) {
                    u32 id_b = c.index_neigh_map[k];
                    vec3 dr  = xyz_a - vec3{st.mxyz[3 * id_b], st.mxyz[3 * id_b + 1], st.mxyz[3 * id_b + 2]};
                    f64 rab2 = dot(dr, dr);
                    f64 h_b  = st.g_h[id_b];
                    if (rab2 > h_a * h_a * Rker2 && rab2 > h_b * h_b * Rker2)
                        continue;
                    f64 rab = std::sqrt(rab2);
                    vec3 vxyz_b{st.g_v[3 * id_b], st.g_v[3 * id_b + 1], st.g_v[3 * id_b + 2]};
                    vec3 v_ab      = vxyz_a - vxyz_b;
                    vec3 r_ab_unit = dr / rab;
                    if (rab < 1e-9)
                        r_ab_unit = {0, 0, 0};
                    f64 abs_v_ab_r_ab = std::fabs(dot(v_ab, r_ab_unit));
                    f64 vsig_a        = 1.0 * cs_a + 2.0 * abs_v_ab_r_ab;
                    vsig_max          = std::fmax(vsig_max, vsig_a);
                }
                st.vsig[id_a] = vsig_max;
            }
        }
    }

    // ---- the step ----------------------------------------------------------------------------
    /// ref: shammodels/sph/src/Solver.cpp:1942-3272
    void evolve_once() {
        f64 t_current = S.time;
        f64 dt        = S.dt;

        point_mass_accrete();
        predictor(dt);
        kill_particles();
        compute_ext_forces_indep_v();
        apply_position_boundary();
        u64 Npart_all = S.total_count();
        S.log.npart   = Npart_all;

        // ref: Solver.cpp:2043-2048
        if (cfg.enable_particle_reordering && S.step_count % cfg.particle_reordering_step_freq == 0)
            S.reorder_particles();

        sph_prestep();

        f64 next_cfl              = 0;
        u32 corrector_iter_cnt    = 0;
        bool need_rerun_corrector = false;
        do {
            if (corrector_iter_cnt == 50)
                throw std::runtime_error("the corrector has made over 50 loops");
            communicate_merge_ghosts_fields();
            for (auto &p : S.patches)
                if (p.pdat.n && cfg.has_alphaAV())
                    S.step[p.id].alpha_updated = p.pdat.alpha_AV;
            if (cfg.has_dtdivv()) {
                if (cfg.combined_dtdiv_divcurlv_compute) {
                    update_dtdivv(true);
                } else {
                    update_divv_curlv(cfg.has_curlv());
                    update_dtdivv(false);
                }
            } else if (cfg.has_divv()) {
                update_divv_curlv(cfg.has_curlv());
            }
            update_artificial_viscosity(dt);
            exchange_alpha_ghosts();
            compute_eos_fields();
            // prepare_corrector (Solver.cpp:1689-1690)
            std::map<u64, std::vector<f64>> old_axyz, old_duint;
            for (auto &p : S.patches) {
                old_axyz[p.id]  = p.pdat.axyz;
                old_duint[p.id] = p.pdat.duint;
            }
            update_derivs();
            // leapfrog corrector (shamrock/src/math/integrators.cpp:88-119), hdt = dt/2
            f64 hdt          = dt / 2;
            f64 max_eps_v_sq = std::numeric_limits<f64>::lowest();
            f64 sum_vsq      = 0;
            for (auto &p : S.patches) {
                auto &d = p.pdat;
                for (u32 i = 0; i < d.n; i++) {
                    vec3 incr{hdt * (d.axyz[3 * i] - old_axyz[p.id][3 * i]),
                              hdt * (d.axyz[3 * i + 1] - old_axyz[p.id][3 * i + 1]),
                              hdt * (d.axyz[3 * i + 2] - old_axyz[p.id][3 * i + 2])};
                    d.vxyz[3 * i]     = d.vxyz[3 * i] + incr.x;
                    d.vxyz[3 * i + 1] = d.vxyz[3 * i + 1] + incr.y;
                    d.vxyz[3 * i + 2] = d.vxyz[3 * i + 2] + incr.z;
                    max_eps_v_sq      = std::fmax(max_eps_v_sq, dot(incr, incr));
                    f64 incu          = hdt * (d.duint[i] - old_duint[p.id][i]);
                    d.uint[i]         = d.uint[i] + incu;
                }
            }
            // Σ v·v : the reduction order is not part of the contract (tree reduction in the
            // reference); accumulate per patch in index order.
            for (auto &p : S.patches) {
                auto &d = p.pdat;
                for (u32 i = 0; i < d.n; i++)
                    sum_vsq += d.vxyz[3 * i] * d.vxyz[3 * i] + d.vxyz[3 * i + 1] * d.vxyz[3 * i + 1]
                               + d.vxyz[3 * i + 2] * d.vxyz[3 * i + 2];
            }
            f64 rank_veps_v = std::sqrt(max_eps_v_sq);
            f64 vmean_sq    = sum_vsq / f64(Npart_all);
            f64 vmean       = std::sqrt(vmean_sq);
            f64 eps_v       = rank_veps_v / vmean;
            if (vmean <= 0)
                eps_v = 0;
            S.log.eps_v = eps_v;
            if (eps_v > 1e-2) {
                need_rerun_corrector = true;
                S.cfl_multiplier     = S.cfl_multiplier / 2;
            } else {
                need_rerun_corrector = false;
            }
            if (!need_rerun_corrector) {
                if (cfg.has_alphaAV())
                    for (auto &p : S.patches)
                        if (p.pdat.n)
                            p.pdat.alpha_AV = S.step[p.id].alpha_updated;
                compute_vsig();
                f64 C_cour  = cfg.cfl_cour * S.cfl_multiplier;
                f64 C_force = cfg.cfl_force * S.cfl_multiplier;
                f64 rank_dt = std::numeric_limits<f64>::infinity();
                for (auto &p : S.patches) {
                    if (p.pdat.n == 0)
                        continue;
                    PatchStep &st = S.step[p.id];
                    st.cfl_dt.assign(st.n, std::numeric_limits<f64>::infinity());
                    for (u32 i = 0; i < st.n; i++) {
                        f64 h_a  = st.g_h[i];
                        f64 dt_c = C_cour * h_a / st.vsig[i];
                        st.cfl_dt[i] = std::fmin(st.cfl_dt[i], dt_c);
                        vec3 a{p.pdat.axyz[3 * i], p.pdat.axyz[3 * i + 1], p.pdat.axyz[3 * i + 2]};
                        f64 dt_f     = C_force * std::sqrt(h_a / length(a));
                        st.cfl_dt[i] = std::fmin(st.cfl_dt[i], dt_f);
                        rank_dt      = std::fmin(rank_dt, st.cfl_dt[i]);
                    }
                }
                next_cfl = rank_dt;
                if (cfg.has_soundspeed_field())
                    for (auto &p : S.patches)
                        if (p.pdat.n)
                            for (u32 i = 0; i < p.pdat.n; i++)
                                p.pdat.soundspeed[i] = S.step[p.id].soundspeed[i];
            }
            corrector_iter_cnt++;
        } while (need_rerun_corrector);
        S.log.corrector_iter = corrector_iter_cnt;
        S.log.next_dt        = next_cfl;
        S.dt                 = next_cfl;
        S.time               = t_current + dt;
        f64 stiff            = cfg.cfl_multiplier_stiffness;
        S.cfl_multiplier     = (S.cfl_multiplier * stiff + 1.) / (stiff + 1.);
        S.step_count++;
    }
};

inline void Solver::evolve_once() {
    if (cfg.kernel == KERNEL_M4) {
        SolverT<KernelM4> s(*this);
        s.evolve_once();
    } else {
        SolverT<KernelM6> s(*this);
        s.evolve_once();
    }
}

} // namespace oracle
