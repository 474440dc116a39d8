// C API of the CPU oracle, for ctypes (tests/, smoke(), bench.py cpu_baseline only).
// TEST INFRASTRUCTURE — never linked into or called from the product library.
#include "sph_step.hpp"
#include <cstdio>
#include <cstring>
#include <omp.h>
#include <string>

using namespace oracle;

namespace {
thread_local std::string g_err;
template<class F>
int guard(F &&f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

struct TreeHandle {
    int bits;
    Tree<u32> t32;
    Tree<u64> t64;
    std::vector<f64> rint;
    ObjectCache cache, leaf_cache;
    std::vector<u32> leaf_owner;
};

template<class T>
int64_t copy_out(const std::vector<T> &v, void *out, int64_t cap_bytes) {
    int64_t nb = int64_t(v.size() * sizeof(T));
    if (out && cap_bytes >= nb && nb > 0)
        std::memcpy(out, v.data(), size_t(nb));
    return nb;
}
int64_t copy_out_vec3(const std::vector<vec3> &v, void *out, int64_t cap_bytes) {
    int64_t nb = int64_t(v.size() * sizeof(vec3));
    if (out && cap_bytes >= nb && nb > 0)
        std::memcpy(out, v.data(), size_t(nb));
    return nb;
}

template<class Tm>
int64_t tree_get(const Tree<Tm> &t, const std::string &n, void *out, int64_t cap) {
    if (n == "sorted_morton") return copy_out(t.sorted_morton, out, cap);
    if (n == "sort_index_map") return copy_out(t.sort_index_map, out, cap);
    if (n == "reduc_index_map") return copy_out(t.reduc_index_map, out, cap);
    if (n == "reduced_morton") return copy_out(t.reduced_morton, out, cap);
    if (n == "lchild_id") return copy_out(t.lchild_id, out, cap);
    if (n == "rchild_id") return copy_out(t.rchild_id, out, cap);
    if (n == "lchild_flag") return copy_out(t.lchild_flag, out, cap);
    if (n == "rchild_flag") return copy_out(t.rchild_flag, out, cap);
    if (n == "endrange") return copy_out(t.endrange, out, cap);
    if (n == "aabb_min") return copy_out_vec3(t.aabb_min, out, cap);
    if (n == "aabb_max") return copy_out_vec3(t.aabb_max, out, cap);
    return -1;
}
int64_t cache_get(const ObjectCache &c, const std::string &n, void *out, int64_t cap) {
    if (n == "cnt_neigh") return copy_out(c.cnt_neigh, out, cap);
    if (n == "scanned_cnt") return copy_out(c.scanned_cnt, out, cap);
    if (n == "index_neigh_map") return copy_out(c.index_neigh_map, out, cap);
    return -1;
}
} // namespace

extern "C" {

const char *oracle_last_error() { return g_err.c_str(); }
int oracle_num_threads() { return omp_get_max_threads(); }
void oracle_set_num_threads(int n) { omp_set_num_threads(n); }

// ---- primitives ---------------------------------------------------------------------------
int oracle_morton_codes(
    int bits, const double *xyz, uint64_t stride_dbl, uint32_t cnt, const double *bmin,
    const double *bmax, uint32_t morton_count, void *out) {
    return guard([&] {
        if (bits == 32)
            morton_code_set_from_positions<u32>(xyz, stride_dbl, cnt, bmin, bmax, morton_count, (u32 *) out);
        else
            morton_code_set_from_positions<u64>(xyz, stride_dbl, cnt, bmin, bmax, morton_count, (u64 *) out);
    });
}
int oracle_sort_by_key(int bits, void *keys, uint32_t *vals, uint32_t len) {
    return guard([&] {
        if (bits == 32)
            sort_by_key_bitonic<u32>((u32 *) keys, vals, len);
        else
            sort_by_key_bitonic<u64>((u64 *) keys, vals, len);
    });
}
/// out_index_map must hold morton_count + 2 entries
int oracle_reduction(
    int bits, const void *sorted, uint32_t morton_count, uint32_t level, uint32_t *out_index_map,
    uint32_t *leaf_count) {
    return guard([&] {
        std::vector<u32> m;
        u32 lc = 0;
        if (bits == 32)
            reduction_alg<u32>((const u32 *) sorted, morton_count, level, m, lc);
        else
            reduction_alg<u64>((const u64 *) sorted, morton_count, level, m, lc);
        std::memcpy(out_index_map, m.data(), m.size() * sizeof(u32));
        *leaf_count = lc;
    });
}
int oracle_karras(
    int bits, const void *codes, uint32_t leaf_count, uint32_t *lchild, uint32_t *rchild,
    uint8_t *lflag, uint8_t *rflag, uint32_t *endrange) {
    return guard([&] {
        if (bits == 32)
            karras_alg<u32>((const u32 *) codes, leaf_count - 1, lchild, rchild, lflag, rflag, endrange);
        else
            karras_alg<u64>((const u64 *) codes, leaf_count - 1, lchild, rchild, lflag, rflag, endrange);
    });
}

// ---- tree handle ----------------------------------------------------------------------------
void *oracle_tree_build(
    int bits, const double *xyz, uint64_t stride_dbl, uint32_t cnt, const double *bmin,
    const double *bmax, uint32_t level, uint32_t morton_count_override) {
    auto *h = new TreeHandle();
    h->bits = bits;
    int rc  = guard([&] {
        if (bits == 32)
            h->t32 = rebuild_from_positions<u32>(xyz, stride_dbl, cnt, bmin, bmax, level, true, morton_count_override);
        else
            h->t64 = rebuild_from_positions<u64>(xyz, stride_dbl, cnt, bmin, bmax, level, true, morton_count_override);
    });
    if (rc) {
        delete h;
        return nullptr;
    }
    return h;
}
void oracle_tree_free(void *h) { delete (TreeHandle *) h; }
void oracle_tree_sizes(void *hh, uint32_t *out4) {
    auto *h = (TreeHandle *) hh;
    if (h->bits == 32) {
        out4[0] = h->t32.obj_cnt; out4[1] = h->t32.morton_count; out4[2] = h->t32.leaf_count; out4[3] = h->t32.int_count;
    } else {
        out4[0] = h->t64.obj_cnt; out4[1] = h->t64.morton_count; out4[2] = h->t64.leaf_count; out4[3] = h->t64.int_count;
    }
}
int64_t oracle_tree_get(void *hh, const char *name, void *out, int64_t cap) {
    auto *h = (TreeHandle *) hh;
    std::string n(name);
    if (n == "rint") return copy_out(h->rint, out, cap);
    if (n.rfind("cache.", 0) == 0) return cache_get(h->cache, n.substr(6), out, cap);
    if (n.rfind("leaf_cache.", 0) == 0) return cache_get(h->leaf_cache, n.substr(11), out, cap);
    if (n == "leaf_owner") return copy_out(h->leaf_owner, out, cap);
    return h->bits == 32 ? tree_get(h->t32, n, out, cap) : tree_get(h->t64, n, out, cap);
}
/// rint[node] = max_h(node) * htol  (Solver.cpp:1322-1356); htol = 1 gives the plain max field
int oracle_tree_field_max(void *hh, const double *field, double htol) {
    auto *h = (TreeHandle *) hh;
    return guard([&] {
        h->rint = h->bits == 32 ? compute_tree_field_max_field(h->t32, field)
                                : compute_tree_field_max_field(h->t64, field);
        for (auto &v : h->rint) v *= htol;
    });
}
int oracle_tree_neigh_cache(
    void *hh, const double *xyz, uint64_t stride_dbl, const double *hpart, uint32_t obj_cnt,
    double Rkern, double htol, int two_stage) {
    auto *h = (TreeHandle *) hh;
    return guard([&] {
        xyzh_view P{xyz, stride_dbl, hpart};
        if (h->bits == 32) {
            h->cache = two_stage ? neighbour_cache_2stages(h->t32, P, obj_cnt, h->rint, Rkern, htol, &h->leaf_cache, &h->leaf_owner)
                                 : neighbour_cache_1stage(h->t32, P, obj_cnt, h->rint, Rkern, htol);
        } else {
            h->cache = two_stage ? neighbour_cache_2stages(h->t64, P, obj_cnt, h->rint, Rkern, htol, &h->leaf_cache, &h->leaf_owner)
                                 : neighbour_cache_1stage(h->t64, P, obj_cnt, h->rint, Rkern, htol);
        }
    });
}
/// The box query of the reference's CLBVHObjectIteratorTests.cpp:60-229: for every object i,
/// all objects whose position lies in [r_i - s, r_i + s] (s < 0: accept everything), in
/// traversal order.  Result stored in cache.*
int oracle_tree_box_query(void *hh, const double *xyz, uint64_t stride_dbl, uint32_t n, double s) {
    auto *h = (TreeHandle *) hh;
    return guard([&] {
        auto run = [&](auto &t) {
            ObjectCache c;
            c.cnt_neigh.resize(n);
            auto pass = [&](bool fill) {
                for (u32 i = 0; i < n; i++) {
                    vec3 r{xyz[i * stride_dbl], xyz[i * stride_dbl + 1], xyz[i * stride_dbl + 2]};
                    vec3 lo = r + vec3{-s, -s, -s}, hi = r + vec3{s, s, s};
                    u32 cnt = fill ? c.scanned_cnt[i] : 0;
                    rtree_for(
                        t,
                        [&](u32, vec3 nmin, vec3 nmax) {
                            if (s < 0) return true;
                            vec3 il = vmax(nmin, lo), ih = vmin(nmax, hi);
                            return ih.x >= il.x && ih.y >= il.y && ih.z >= il.z;
                        },
                        [&](u32 leaf) {
                            t.for_each_in_leaf_cell(leaf - t.int_count, [&](u32 id) {
                                vec3 r2{xyz[id * stride_dbl], xyz[id * stride_dbl + 1], xyz[id * stride_dbl + 2]};
                                vec3 il = vmax(r2, lo), ih = vmin(r2, hi);
                                bool in = (s < 0) || (ih.x >= il.x && ih.y >= il.y && ih.z >= il.z);
                                if (in) {
                                    if (fill) c.index_neigh_map[cnt] = id;
                                    cnt++;
                                }
                            });
                        });
                    if (!fill) c.cnt_neigh[i] = cnt;
                }
            };
            pass(false);
            prepare_object_cache(c);
            pass(true);
            h->cache = std::move(c);
        };
        if (h->bits == 32) run(h->t32); else run(h->t64);
    });
}

// ---- SPH pieces on raw arrays -----------------------------------------------------------------
int oracle_h_iterate(
    int kernel, const uint32_t *cnt, const uint32_t *scanned, const uint32_t *idx, uint64_t nidx,
    const double *xyz, uint64_t stride_dbl, uint32_t n, const double *h_old, double *h_new,
    double *eps, double pmass, double h_evol_max, double h_evol_iter_max) {
    return guard([&] {
        ObjectCache c;
        c.cnt_neigh.assign(cnt, cnt + n);
        c.scanned_cnt.assign(scanned, scanned + n);
        c.index_neigh_map.assign(idx, idx + nidx);
        if (kernel == KERNEL_M4)
            iterate_smoothing_length_density<KernelM4>(c, xyz, stride_dbl, n, h_old, h_new, eps, pmass, h_evol_max, h_evol_iter_max);
        else
            iterate_smoothing_length_density<KernelM6>(c, xyz, stride_dbl, n, h_old, h_new, eps, pmass, h_evol_max, h_evol_iter_max);
    });
}
/// kernel functions (sphkernelsTests.cpp identities): which = 0 f, 1 df, 2 W_3d, 3 dW_3d, 4 dhW_3d
double oracle_kernel_eval(int kernel, int which, double a, double b) {
    if (kernel == KERNEL_M4) {
        switch (which) {
        case 0: return KernelM4::f(a);
        case 1: return KernelM4::df(a);
        case 2: return SPHKernel<KernelM4>::W_3d(a, b);
        case 3: return SPHKernel<KernelM4>::dW_3d(a, b);
        default: return SPHKernel<KernelM4>::dhW_3d(a, b);
        }
    }
    switch (which) {
    case 0: return KernelM6::f(a);
    case 1: return KernelM6::df(a);
    case 2: return SPHKernel<KernelM6>::W_3d(a, b);
    case 3: return SPHKernel<KernelM6>::dW_3d(a, b);
    default: return SPHKernel<KernelM6>::dhW_3d(a, b);
    }
}

// ---- solver handle ------------------------------------------------------------------------------
void *oracle_solver_create() { return new Solver(); }
void oracle_solver_free(void *s) { delete (Solver *) s; }
/// cfg is passed as "key=value;key=value" to keep the ABI trivial
int oracle_solver_configure(void *ss, const char *kv) {
    auto *S = (Solver *) ss;
    return guard([&] {
        std::string str(kv);
        size_t pos = 0;
        auto &c    = S->cfg;
        while (pos < str.size()) {
            size_t e = str.find(';', pos);
            if (e == std::string::npos) e = str.size();
            std::string item = str.substr(pos, e - pos);
            pos              = e + 1;
            if (item.empty()) continue;
            size_t q = item.find('=');
            std::string k = item.substr(0, q);
            double v      = std::stod(item.substr(q + 1));
            if (k == "kernel") c.kernel = int(v);
            else if (k == "gpart_mass") c.gpart_mass = v;
            else if (k == "eos") c.eos = int(v);
            else if (k == "gamma") c.gamma = v;
            else if (k == "cs0") c.cs0 = v;
            else if (k == "eos_q") c.eos_q = v;
            else if (k == "eos_r0") c.eos_r0 = v;
            else if (k == "av") c.av = int(v);
            else if (k == "alpha_u") c.alpha_u = v;
            else if (k == "alpha_AV") c.alpha_AV = v;
            else if (k == "beta_AV") c.beta_AV = v;
            else if (k == "alpha_min") c.alpha_min = v;
            else if (k == "alpha_max") c.alpha_max = v;
            else if (k == "sigma_decay") c.sigma_decay = v;
            else if (k == "bc") c.bc = int(v);
            else if (k == "cfl_cour") c.cfl_cour = v;
            else if (k == "cfl_force") c.cfl_force = v;
            else if (k == "cfl_multiplier_stiffness") c.cfl_multiplier_stiffness = v;
            else if (k == "htol_up_coarse_cycle") c.htol_up_coarse_cycle = v;
            else if (k == "htol_up_fine_cycle") c.htol_up_fine_cycle = v;
            else if (k == "epsilon_h") c.epsilon_h = v;
            else if (k == "h_iter_per_subcycles") c.h_iter_per_subcycles = u32(v);
            else if (k == "h_max_subcycles_count") c.h_max_subcycles_count = u32(v);
            else if (k == "tree_reduction_level") c.tree_reduction_level = u32(v);
            else if (k == "use_two_stage_search") c.use_two_stage_search = int(v);
            else if (k == "combined_dtdiv_divcurlv_compute") c.combined_dtdiv_divcurlv_compute = int(v);
            else if (k == "has_point_mass") c.has_point_mass = int(v);
            else if (k == "pm_mass") c.pm_mass = v;
            else if (k == "pm_racc") c.pm_racc = v;
            else if (k == "constant_G") c.constant_G = v;
            else if (k == "enable_particle_reordering") c.enable_particle_reordering = int(v);
            else if (k == "particle_reordering_step_freq") {
                if (v < 1) throw std::invalid_argument("particle_reordering_step_freq cannot be zero");
                c.particle_reordering_step_freq = u64(v);
            }
            else if (k == "time") S->time = v;
            else if (k == "dt") S->dt = v;
            else if (k == "cfl_multiplier") S->cfl_multiplier = v;
            else throw std::invalid_argument("unknown config key: " + k);
        }
    });
}
int oracle_solver_add_kill_sphere(void *ss, const double *c, double r) {
    auto *S = (Solver *) ss;
    return guard([&] {
        int k = S->cfg.n_kill_spheres;
        if (k >= 4) throw std::runtime_error("too many kill spheres");
        for (int d = 0; d < 3; d++) S->cfg.kill_center[k][d] = c[d];
        S->cfg.kill_radius[k] = r;
        S->cfg.n_kill_spheres++;
    });
}
int oracle_solver_set_box(void *ss, const double *bmin, const double *bmax, uint32_t nx, uint32_t ny, uint32_t nz) {
    auto *S = (Solver *) ss;
    return guard([&] {
        for (int d = 0; d < 3; d++) { S->box_min[d] = bmin[d]; S->box_max[d] = bmax[d]; }
        S->init_patch_grid(nx, ny, nz);
    });
}
uint32_t oracle_solver_patch_count(void *ss) { return u32(((Solver *) ss)->patches.size()); }
uint32_t oracle_solver_patch_size(void *ss, uint32_t ip) { return ((Solver *) ss)->patches[ip].pdat.n; }
/// append particles (distributed to patches by position); fields xyz,vxyz (3n), hpart,uint (n);
/// other fields start at 0 (alpha_AV at alpha_min is set by the caller through set_field)
int oracle_solver_push_particles(void *ss, uint32_t n, const double *xyz, const double *vxyz, const double *h, const double *u) {
    auto *S = (Solver *) ss;
    return guard([&] {
        for (u32 i = 0; i < n; i++) {
            int own = S->patch_owner(&xyz[3 * i]);
            if (own < 0) {
                if (S->patches.size() == 1) own = 0;
                else throw std::runtime_error("particle outside of the simulation box");
            }
            auto &d = S->patches[own].pdat;
            for (int c = 0; c < 3; c++) {
                d.xyz.push_back(xyz[3 * i + c]);
                d.vxyz.push_back(vxyz ? vxyz[3 * i + c] : 0.);
                d.axyz.push_back(0.); d.axyz_ext.push_back(0.); d.curlv.push_back(0.);
            }
            d.hpart.push_back(h[i]);
            d.uint.push_back(u ? u[i] : 0.);
            d.duint.push_back(0.); d.alpha_AV.push_back(0.); d.divv.push_back(0.);
            d.dtdivv.push_back(0.); d.soundspeed.push_back(0.);
            d.n++;
        }
    });
}
static std::vector<f64> *field_ptr(PatchData &d, const std::string &n, int &nv) {
    nv = 1;
    if (n == "xyz") { nv = 3; return &d.xyz; }
    if (n == "vxyz") { nv = 3; return &d.vxyz; }
    if (n == "axyz") { nv = 3; return &d.axyz; }
    if (n == "axyz_ext") { nv = 3; return &d.axyz_ext; }
    if (n == "curlv") { nv = 3; return &d.curlv; }
    if (n == "hpart") return &d.hpart;
    if (n == "uint") return &d.uint;
    if (n == "duint") return &d.duint;
    if (n == "alpha_AV") return &d.alpha_AV;
    if (n == "divv") return &d.divv;
    if (n == "dtdivv") return &d.dtdivv;
    if (n == "soundspeed") return &d.soundspeed;
    return nullptr;
}
int oracle_solver_set_field(void *ss, uint32_t ip, const char *name, const double *in) {
    auto *S = (Solver *) ss;
    return guard([&] {
        int nv;
        auto *f = field_ptr(S->patches.at(ip).pdat, name, nv);
        if (!f) throw std::invalid_argument(std::string("unknown field ") + name);
        std::memcpy(f->data(), in, f->size() * sizeof(f64));
    });
}
/// names: main fields; step data "step.<name>" (mxyz, mh, rint, omega, pressure, soundspeed,
/// g_h, g_u, g_v, g_a, g_omega, g_alpha, vsig, cfl_dt, alpha_updated), "tree.<name>",
/// "cache.<name>".  Returns the byte size (copying when cap is large enough) or -1.
int64_t oracle_solver_get(void *ss, uint32_t ip, const char *name, void *out, int64_t cap) {
    auto *S = (Solver *) ss;
    std::string n(name);
    if (ip >= S->patches.size()) return -1;
    auto &p = S->patches[ip];
    int nv;
    if (auto *f = field_ptr(p.pdat, n, nv)) return copy_out(*f, out, cap);
    auto it = S->step.find(p.id);
    if (it == S->step.end()) return -1;
    PatchStep &st = it->second;
    if (n.rfind("tree.", 0) == 0) return tree_get(st.tree, n.substr(5), out, cap);
    if (n.rfind("cache.", 0) == 0) return cache_get(st.cache, n.substr(6), out, cap);
    if (n == "step.mxyz") return copy_out(st.mxyz, out, cap);
    if (n == "step.mh") return copy_out(st.mh, out, cap);
    if (n == "step.rint") return copy_out(st.rint, out, cap);
    if (n == "step.omega") return copy_out(st.omega, out, cap);
    if (n == "step.pressure") return copy_out(st.pressure, out, cap);
    if (n == "step.soundspeed") return copy_out(st.soundspeed, out, cap);
    if (n == "step.g_h") return copy_out(st.g_h, out, cap);
    if (n == "step.g_u") return copy_out(st.g_u, out, cap);
    if (n == "step.g_v") return copy_out(st.g_v, out, cap);
    if (n == "step.g_a") return copy_out(st.g_a, out, cap);
    if (n == "step.g_omega") return copy_out(st.g_omega, out, cap);
    if (n == "step.g_alpha") return copy_out(st.g_alpha, out, cap);
    if (n == "step.vsig") return copy_out(st.vsig, out, cap);
    if (n == "step.cfl_dt") return copy_out(st.cfl_dt, out, cap);
    if (n == "step.alpha_updated") return copy_out(st.alpha_updated, out, cap);
    return -1;
}
/// modules::ParticleReordering::reorder_particles (the call SPHSetup::apply_setup makes, SPHSetup.cpp:202-205)
int oracle_solver_reorder_particles(void *ss) {
    auto *S = (Solver *) ss;
    return guard([&] { S->reorder_particles(); });
}
int oracle_solver_split_patch(void *ss, uint32_t ip) {
    auto *S = (Solver *) ss;
    return guard([&] { S->split_patch(ip); });
}
int oracle_solver_merge_patches(void *ss, uint32_t ip0) {
    auto *S = (Solver *) ss;
    return guard([&] { S->merge_patches(ip0); });
}
uint64_t oracle_solver_patch_id(void *ss, uint32_t ip) { return ((Solver *) ss)->patches.at(ip).id; }
int oracle_solver_evolve_once(void *ss) {
    auto *S = (Solver *) ss;
    return guard([&] { S->evolve_once(); });
}
/// out: time, dt(next), cfl_multiplier, eps_v, h_subcycles, h_iters_last, corrector_iter, npart
void oracle_solver_state(void *ss, double *out8) {
    auto *S = (Solver *) ss;
    out8[0] = S->time; out8[1] = S->dt; out8[2] = S->cfl_multiplier; out8[3] = S->log.eps_v;
    out8[4] = S->log.h_subcycles; out8[5] = S->log.h_iters_last; out8[6] = S->log.corrector_iter;
    out8[7] = double(S->log.npart);
}
void oracle_solver_set_dt(void *ss, double dt) { ((Solver *) ss)->dt = dt; }
void oracle_solver_set_gpart_mass(void *ss, double m) { ((Solver *) ss)->cfg.gpart_mass = m; }

} // extern "C"
