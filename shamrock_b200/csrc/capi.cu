// capi.cu — the extern "C" boundary declared in include/shamb200.h.
#include "solver.cuh"
#include "load_balance.hpp"
#include <cstring>
#include <mutex>
#include <set>

using namespace sb;


namespace {
thread_local std::string g_err;

template<class F>
int guard(F &&f) {
    try {
        f();
        return SHAMB200_OK;
    } catch (const CudaError &e) {
        g_err = e.what();
        return SHAMB200_ERR_CUDA;
    } catch (const std::overflow_error &e) {
        g_err = e.what();
        return SHAMB200_ERR_OVERFLOW;
    } catch (const std::invalid_argument &e) {
        g_err = e.what();
        return SHAMB200_ERR_INVALID;
    } catch (const std::exception &e) {
        g_err = e.what();
        return SHAMB200_ERR_RUNTIME;
    }
}

void fill_tree_view(const TreeBuffers &t, shamb200_tree *out) {
    out->obj_cnt      = t.M;
    out->morton_count = t.P2;
    out->leaf_count   = t.L;
    out->int_count    = t.I;
    for (int d = 0; d < 3; d++) {
        out->bmin[d] = t.bmin[d];
        out->bmax[d] = t.bmax[d];
    }
    out->d_sorted_morton   = t.morton.p;
    out->d_sort_index_map  = t.index_map.p;
    out->d_reduc_index_map = t.reduc_index_map.p;
    out->d_reduced_morton  = t.reduced_morton.p;
    out->d_lchild_id       = t.lchild.p;
    out->d_rchild_id       = t.rchild.p;
    out->d_endrange        = t.endrange.p;
    out->d_lchild_flag     = t.lflag.p;
    out->d_rchild_flag     = t.rflag.p;
    out->d_aabb_min        = t.aabb_min.p;
    out->d_aabb_max        = t.aabb_max.p;
}
} // namespace

namespace sb {
void set_last_error(const char *msg) { g_err = msg; }
} // namespace sb

// Live handles.  A binding with garbage collection finalises a model and its context in any order (and may
// try to finalise a model whose context is already gone): destroying a context first destroys the models that
// run on it, destroying a handle that is not live is a no-op, and using one is SHAMB200_ERR_INVALID.
namespace {
std::mutex g_handles_mu;
std::set<shamb200_model *> g_live_models;
std::set<shamb200_ctx *> g_live_ctxs;

bool model_is_live(shamb200_model *m) {
    std::lock_guard<std::mutex> lk(g_handles_mu);
    return m && g_live_models.count(m);
}
void need_live(shamb200_model *m) {
    if (!model_is_live(m))
        throw std::invalid_argument("stale model handle (the model or its context has been destroyed)");
    pool_set_stream(m->m.ctx->stream);
}
void need_live(shamb200_ctx *c) {
    std::lock_guard<std::mutex> lk(g_handles_mu);
    if (!c || !g_live_ctxs.count(c))
        throw std::invalid_argument("stale context handle");
    pool_set_stream(c->c.stream);
}
} // namespace
namespace sb {
void require_live(shamb200_ctx *c) { need_live(c); }
} // namespace sb
namespace {
/// the model is live and owned by the caller (already taken out of the registry)
void destroy_model_now(shamb200_model *m) {
    cudaSetDevice(m->m.ctx->device);
    pool_set_stream(m->m.ctx->stream);
    cudaStreamSynchronize(m->m.ctx->stream);
    comm_destroy(m->m);
    delete m;
    cudaGetLastError(); // a failure while tearing down must not surface in a later launch check
}
} // namespace

extern "C" {

const char *shamb200_last_error(void) { return g_err.c_str(); }
const char *shamb200_build_info(void) {
    return "shamb200 sm_100a fp=strict(no-fma, bit-exact with the oracle)+fast(fma), selected by fp_mode";
}
uint64_t shamb200_launch_count(void) { return g_launch_count; }
void shamb200_reset_launch_count(void) { g_launch_count = 0; }

int shamb200_ctx_create(int device, void *cuda_stream, shamb200_ctx **out) {
    return guard([&] {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw CudaError(
                std::string("no CUDA device available (") + cudaGetErrorString(e)
                + "): shamb200 has no CPU fallback");
        if (device < 0 || device >= ndev)
            throw std::invalid_argument("invalid device index");
        SB_CUDA_CHECK(cudaSetDevice(device));
        auto *c     = new shamb200_ctx();
        c->c.device = device;
        if (cuda_stream) {
            c->c.stream     = (cudaStream_t) cuda_stream;
            c->c.own_stream = false;
        } else {
            SB_CUDA_CHECK(cudaStreamCreateWithFlags(&c->c.stream, cudaStreamNonBlocking));
            c->c.own_stream = true;
        }
        {
            std::lock_guard<std::mutex> lk(g_handles_mu);
            g_live_ctxs.insert(c);
        }
        *out = c;
    });
}
int shamb200_ctx_destroy(shamb200_ctx *ctx) {
    return guard([&] {
        std::vector<shamb200_model *> mine;
        {
            std::lock_guard<std::mutex> lk(g_handles_mu);
            if (!ctx || !g_live_ctxs.erase(ctx))
                return; // not a live context (already destroyed)
            for (auto it = g_live_models.begin(); it != g_live_models.end();) {
                if ((*it)->m.ctx == &ctx->c) {
                    mine.push_back(*it);
                    it = g_live_models.erase(it);
                } else {
                    ++it;
                }
            }
        }
        for (auto *m : mine) // a model never outlives the stream it runs on
            destroy_model_now(m);
        cudaSetDevice(ctx->c.device);
        cudaStream_t st = ctx->c.stream;
        const bool own  = ctx->c.own_stream;
        pool_set_stream(st);
        cudaStreamSynchronize(st);
        delete ctx; // its buffers go back to the pool while the stream still exists
        cudaStreamSynchronize(st);
        pool_set_stream(nullptr);
        if (own)
            cudaStreamDestroy(st);
        cudaGetLastError(); // a failure while tearing down must not surface in a later launch check
    });
}
void *shamb200_ctx_stream(shamb200_ctx *ctx) {
    std::lock_guard<std::mutex> lk(g_handles_mu);
    return ctx && g_live_ctxs.count(ctx) ? (void *) ctx->c.stream : nullptr;
}
int shamb200_ctx_synchronize(shamb200_ctx *ctx) {
    return guard([&] {
        need_live(ctx);
        SB_CUDA_CHECK(cudaStreamSynchronize(ctx->c.stream));
    });
}

int shamb200_tree_build(
    shamb200_ctx *ctx, const double *d_xyz, size_t stride_dbl, uint32_t obj_cnt, const double bmin[3],
    const double bmax[3], uint32_t reduction_level, int sort_mode, shamb200_tree *out) {
    return guard([&] {
        need_live(ctx);
        SB_CUDA_CHECK(cudaSetDevice(ctx->c.device));
        tree_build(ctx->c.stream, ctx->c.api_tree, d_xyz, stride_dbl, obj_cnt, bmin, bmax, false, reduction_level, sort_mode);
        fill_tree_view(ctx->c.api_tree, out);
    });
}
int shamb200_tree_build_auto_bbox(
    shamb200_ctx *ctx, const double *d_xyz, size_t stride_dbl, uint32_t obj_cnt, uint32_t reduction_level,
    int sort_mode, shamb200_tree *out) {
    return guard([&] {
        need_live(ctx);
        SB_CUDA_CHECK(cudaSetDevice(ctx->c.device));
        tree_build(ctx->c.stream, ctx->c.api_tree, d_xyz, stride_dbl, obj_cnt, nullptr, nullptr, true, reduction_level, sort_mode);
        fill_tree_view(ctx->c.api_tree, out);
    });
}
int shamb200_tree_field_max(
    shamb200_ctx *ctx, const shamb200_tree *tree, const double *d_field, double scale, double *d_out) {
    return guard([&] {
        need_live(ctx);
        if (tree->d_sort_index_map != ctx->c.api_tree.index_map.p)
            throw std::invalid_argument("the tree view is stale (not the last tree built by this context)");
        tree_field_max(ctx->c.stream, ctx->c.api_tree, d_field, scale, d_out, 1);
    });
}

int shamb200_neigh_cache_build(
    shamb200_ctx *ctx, const shamb200_tree *tree, const double *d_xyz, size_t stride_dbl, const double *d_hpart,
    const double *d_rint, uint32_t obj_cnt, double Rkern, double h_tolerance, int two_stage, shamb200_csr *out) {
    return guard([&] {
        need_live(ctx);
        if (tree->d_sort_index_map != ctx->c.api_tree.index_map.p)
            throw std::invalid_argument("the tree view is stale (not the last tree built by this context)");
        if (two_stage) { // the B200 search (neigh2.cu), exported in the reference's ObjectCache layout
            SearchBuffers &sb = ctx->c.api_srch;
            search_prepare_sorted_strided(ctx->c.stream, ctx->c.api_tree, sb, d_xyz, stride_dbl, d_hpart, 1, obj_cnt);
            search_build(ctx->c.stream, ctx->c.api_tree, sb, d_rint, Rkern, h_tolerance);
            export_object_cache(ctx->c.stream, ctx->c.api_tree, sb);
            out->obj_cnt           = sb.N;
            out->sum_neigh_cnt     = u32(sb.K);
            out->d_cnt_neigh       = sb.x_cnt.p;
            out->d_scanned_cnt     = sb.x_scanned.p;
            out->d_index_neigh_map = sb.x_list.p;
            return;
        }
        neigh_cache_build(
            ctx->c.stream, ctx->c.api_tree, ctx->c.api_nb, d_xyz, stride_dbl, d_hpart, d_rint, obj_cnt, Rkern,
            h_tolerance, false, 1);
        out->obj_cnt           = ctx->c.api_nb.N;
        out->sum_neigh_cnt     = ctx->c.api_nb.K;
        out->d_cnt_neigh       = ctx->c.api_nb.cnt.p;
        out->d_scanned_cnt     = ctx->c.api_nb.scanned.p;
        out->d_index_neigh_map = ctx->c.api_nb.list.p;
    });
}

int shamb200_neigh_cache_stats(shamb200_ctx *ctx, uint64_t out[6]) {
    return guard([&] {
        need_live(ctx);
        const SearchBuffers &sb = ctx->c.api_srch;
        out[0] = sb.K;
        out[1] = sb.pair_tests;
        out[2] = sb.attempts_last;
        out[3] = sb.over_groups_last;
        out[4] = sb.frontier_cap;
        out[5] = sb.gcand.cap;
    });
}

static void ctx_reset_red(Ctx &c) {
    c.red.ensure(8);
    c.h_red.ensure(8);
    c.h_red.p[0] = 0;
    c.h_red.p[1] = 0xFFFFFFFFFFFFFFFFull;
    SB_CUDA_CHECK(cudaMemcpyAsync(c.red.p, c.h_red.p, 2 * sizeof(u64), cudaMemcpyHostToDevice, c.stream));
}

int shamb200_h_iterate(
    shamb200_ctx *ctx, int kernel, const shamb200_csr *csr, const double *d_xyz, size_t stride_dbl,
    const double *d_h_old, double *d_h_new, double *d_eps, double gpart_mass, double h_evol_max,
    double h_evol_iter_max) {
    return guard([&] {
        need_live(ctx);
        ctx_reset_red(ctx->c);
        CsrView c{csr->d_cnt_neigh, csr->d_scanned_cnt, csr->d_index_neigh_map, csr->obj_cnt};
        h_iterate(ctx->c.stream, kernel, c, d_xyz, stride_dbl, nullptr, 0, d_h_old, d_h_new, d_eps, gpart_mass, h_evol_max, h_evol_iter_max, ctx->c.red.p);
    });
}
int shamb200_h_iterate_loop(
    shamb200_ctx *ctx, int kernel, const shamb200_csr *csr, const double *d_xyz, size_t stride_dbl,
    const double *d_h_old, double *d_h_new, double *d_eps, double gpart_mass, double h_evol_max,
    double h_evol_iter_max, double epsilon_h, uint32_t max_sweeps, double out3[3]) {
    return guard([&] {
        need_live(ctx);
        CsrView c{csr->d_cnt_neigh, csr->d_scanned_cnt, csr->d_index_neigh_map, csr->obj_cnt};
        f64 mx = std::numeric_limits<f64>::max(), mn = -1;
        u32 it = 0;
        for (; it < max_sweeps; it++) {
            ctx_reset_red(ctx->c);
            h_iterate(ctx->c.stream, kernel, c, d_xyz, stride_dbl, nullptr, 0, d_h_old, d_h_new, d_eps, gpart_mass, h_evol_max, h_evol_iter_max, ctx->c.red.p);
            SB_CUDA_CHECK(cudaMemcpyAsync(ctx->c.h_red.p + 2, ctx->c.red.p, 2 * sizeof(u64), cudaMemcpyDeviceToHost, ctx->c.stream));
            SB_CUDA_CHECK(cudaStreamSynchronize(ctx->c.stream));
            mx = ordered_to_f64(ctx->c.h_red.p[2]);
            mn = ordered_to_f64(ctx->c.h_red.p[3]);
            if (mx < epsilon_h)
                break;
        }
        out3[0] = mx;
        out3[1] = mn;
        out3[2] = f64(it);
    });
}
int shamb200_compute_omega(
    shamb200_ctx *ctx, int kernel, const shamb200_csr *csr, const double *d_xyz, size_t stride_dbl,
    const double *d_hpart, double *d_omega, double gpart_mass) {
    return guard([&] {
        need_live(ctx);
        CsrView c{csr->d_cnt_neigh, csr->d_scanned_cnt, csr->d_index_neigh_map, csr->obj_cnt};
        compute_omega(ctx->c.stream, kernel, c, d_xyz, stride_dbl, nullptr, 0, d_hpart, d_omega, gpart_mass);
    });
}

int shamb200_microbench(shamb200_ctx *ctx, int what, double *out) {
    return guard([&] {
        need_live(ctx);
        SB_CUDA_CHECK(cudaSetDevice(ctx->c.device));
        *out = microbench(ctx->c, what);
    });
}

// ---- planning (host only) -------------------------------------------------------------------------
int shamb200_plan_patch_grid(
    const double bmin[3], const double bmax[3], uint32_t nx, uint32_t ny, uint32_t nz, int world_size,
    double *boxes, int32_t *owner) {
    return guard([&] {
        auto g = plan_patch_grid(bmin, bmax, nx, ny, nz, world_size);
        for (size_t k = 0; k < g.size(); k++) {
            for (int d = 0; d < 3; d++) {
                boxes[6 * k + d]     = g[k].lo[d];
                boxes[6 * k + 3 + d] = g[k].hi[d];
            }
            owner[k] = g[k].owner;
        }
    });
}
uint64_t shamb200_hilbert_index(uint64_t x, uint64_t y, uint64_t z) { return hilbert_index_3d(x, y, z); }
int shamb200_plan_load_balance(
    uint32_t npatch, const uint64_t *coord_min, const uint64_t *load, int world_size, int32_t *owner, int *strategy) {
    return guard([&] {
        if ((npatch && (!coord_min || !load || !owner)))
            throw std::invalid_argument("load balance: null argument");
        std::vector<uint64_t> c(coord_min, coord_min + size_t(3) * npatch), l(load, load + npatch);
        auto o = hilbert_load_balance(c, l, world_size, strategy);
        std::copy(o.begin(), o.end(), owner);
    });
}
int shamb200_model_set_patch_owners(shamb200_model *m, uint32_t npatch, const int32_t *owner) {
    return guard([&] {
        need_live(m);
        Model &M = m->m;
        if (npatch != M.patches.size())
            throw std::invalid_argument("set_patch_owners: one owner per patch of the grid");
        for (auto &p : M.patches)
            if (p.f.n)
                throw std::invalid_argument("set_patch_owners: particles have already been pushed");
        for (uint32_t k = 0; k < npatch; k++) {
            if (owner[k] < 0 || owner[k] >= M.world)
                throw std::invalid_argument("set_patch_owners: owner outside the world");
            M.patches[k].owner = owner[k];
        }
    });
}
int shamb200_model_patch_coords(shamb200_model *m, uint32_t npatch, uint64_t *coord_min) {
    return guard([&] {
        need_live(m);
        Model &M = m->m;
        if (npatch != M.patches.size() || !coord_min)
            throw std::invalid_argument("patch_coords: one entry per patch of the grid");
        for (uint32_t k = 0; k < npatch; k++)
            for (int d = 0; d < 3; d++)
                coord_min[3 * k + d] = M.patches[k].cmin[d];
    });
}
int shamb200_plan_interfaces(
    uint32_t npatch, const double *boxes, const double bmin[3], const double bmax[3], int periodic,
    const double *interact_r, const uint32_t *pcount, uint32_t cap, shamb200_iface *out, uint32_t *n_found) {
    return guard([&] {
        std::vector<PatchBox> pb(npatch);
        for (uint32_t k = 0; k < npatch; k++) {
            pb[k].id = k;
            for (int d = 0; d < 3; d++) {
                pb[k].lo[d] = boxes[6 * k + d];
                pb[k].hi[d] = boxes[6 * k + 3 + d];
            }
        }
        auto c = plan_interfaces(
            pb, bmin, bmax, periodic != 0, std::vector<double>(interact_r, interact_r + npatch),
            std::vector<uint32_t>(pcount, pcount + npatch));
        *n_found = uint32_t(c.size());
        for (size_t q = 0; q < c.size() && q < cap; q++) {
            out[q].sender   = c[q].sender;
            out[q].receiver = c[q].receiver;
            for (int d = 0; d < 3; d++) {
                out[q].ioff[d]   = c[q].ioff[d];
                out[q].offset[d] = c[q].offset[d];
                out[q].cut_lo[d] = c[q].cut_lo[d];
                out[q].cut_hi[d] = c[q].cut_hi[d];
            }
        }
    });
}

// ---- model --------------------------------------------------------------------------------------
void shamb200_solver_config_default(shamb200_solver_config *cfg) {
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->kernel     = SHAMB200_KERNEL_M4;
    cfg->eos        = SHAMB200_EOS_ADIABATIC;
    cfg->av         = SHAMB200_AV_CONSTANT;
    cfg->bc         = SHAMB200_BC_FREE;
    cfg->gpart_mass = 0;
    cfg->gamma      = 5. / 3.;
    cfg->cs0        = 1;
    cfg->eos_q      = 0;
    cfg->eos_r0     = 1;
    cfg->alpha_u    = 1;
    cfg->alpha_AV   = 1;
    cfg->beta_AV    = 2;
    cfg->alpha_min  = 0.1;
    cfg->alpha_max  = 1;
    cfg->sigma_decay = 0.1;
    cfg->cfl_cour    = 0.3;
    cfg->cfl_force   = 0.25;
    cfg->cfl_multiplier_stiffness = 2;
    cfg->htol_up_coarse_cycle     = 1.1;
    cfg->htol_up_fine_cycle       = 1.1;
    cfg->epsilon_h                = 1e-6;
    cfg->h_iter_per_subcycles     = 50;
    cfg->h_max_subcycles_count    = 100;
    cfg->tree_reduction_level     = 3;
    cfg->use_two_stage_search     = 1;
    cfg->sort_mode                = SHAMB200_SORT_BITONIC;
    cfg->constant_G               = 1;
    cfg->enable_particle_reordering    = 0;
    cfg->particle_reordering_step_freq = 1000;
}

/// one validation for create and set_config: enum fields in range, array bounds of the config respected
static void validate_config(const shamb200_solver_config *cfg) {
    if (!cfg)
        throw std::invalid_argument("null solver config");
    auto in = [](int v, int lo, int hi) { return v >= lo && v <= hi; };
    if (!in(cfg->kernel, SHAMB200_KERNEL_M4, SHAMB200_KERNEL_M6))
        throw std::invalid_argument("solver config: unknown kernel");
    if (!in(cfg->eos, SHAMB200_EOS_ADIABATIC, SHAMB200_EOS_LOCALLY_ISOTHERMAL_LP07))
        throw std::invalid_argument("solver config: unknown eos");
    if (!in(cfg->av, SHAMB200_AV_NONE, SHAMB200_AV_CONSTANT_DISC))
        throw std::invalid_argument("solver config: unknown artificial viscosity");
    if (!in(cfg->bc, SHAMB200_BC_FREE, SHAMB200_BC_PERIODIC))
        throw std::invalid_argument("solver config: unknown boundary condition");
    if (!in(cfg->fp_mode, SHAMB200_FP_STRICT, SHAMB200_FP_FAST))
        throw std::invalid_argument("solver config: unknown fp_mode");
    if (!in(cfg->sort_mode, SHAMB200_SORT_BITONIC, SHAMB200_SORT_RADIX))
        throw std::invalid_argument("solver config: unknown sort_mode");
    if (cfg->n_kill_spheres < 0 || cfg->n_kill_spheres > 4)
        throw std::invalid_argument("solver config: n_kill_spheres must be in [0, 4]");
    if (cfg->enable_particle_reordering && cfg->particle_reordering_step_freq == 0)
        throw std::invalid_argument("solver config: particle_reordering_step_freq must be > 0 when reordering is enabled");
}

int shamb200_model_create(shamb200_ctx *ctx, const shamb200_solver_config *cfg, shamb200_model **out) {
    return guard([&] {
        need_live(ctx);
        validate_config(cfg);
        auto *m = new shamb200_model(&ctx->c, *cfg);
        {
            std::lock_guard<std::mutex> lk(g_handles_mu);
            g_live_models.insert(m);
        }
        *out = m;
    });
}
int shamb200_model_destroy(shamb200_model *m) {
    return guard([&] {
        {
            std::lock_guard<std::mutex> lk(g_handles_mu);
            if (!m || !g_live_models.erase(m))
                return; // not a live model (already destroyed, possibly with its context)
        }
        destroy_model_now(m);
    });
}
int shamb200_model_set_config(shamb200_model *m, const shamb200_solver_config *cfg) {
    return guard([&] {
        need_live(m);
        validate_config(cfg);
        m->m.cfg = *cfg;
    });
}
int shamb200_nccl_unique_id(void *out128) {
    int rc = guard([&] {
        if (!out128)
            throw std::invalid_argument("null unique-id buffer");
        comm_unique_id(out128);
    });
    return rc == SHAMB200_ERR_RUNTIME ? SHAMB200_ERR_NCCL : rc;
}
int shamb200_model_init_comm(shamb200_model *m, int rank, int world_size, const void *nccl_id128) {
    int rc = guard([&] {
        need_live(m);
        if (world_size > 1 && !nccl_id128)
            throw std::invalid_argument("null NCCL unique id");
        comm_init(m->m, rank, world_size, nccl_id128);
    });
    return rc == SHAMB200_ERR_RUNTIME ? SHAMB200_ERR_NCCL : rc;
}
/* kept for old bindings: the message is the one of shamb200_last_error */
const char *shamb200_comm_last_error(void) { return g_err.c_str(); }
int shamb200_model_set_box(shamb200_model *m, const double bmin[3], const double bmax[3], uint32_t nx, uint32_t ny, uint32_t nz) {
    return guard([&] {
        need_live(m);
        m->m.set_box(bmin, bmax, nx, ny, nz);
    });
}
int shamb200_model_push_particles(shamb200_model *m, uint64_t n, const double *xyz, const double *vxyz, const double *hpart, const double *uint_) {
    return guard([&] {
        need_live(m);
        m->m.push_particles(n, xyz, vxyz, hpart, uint_);
    });
}
int shamb200_model_dump(shamb200_model *m, const char *fname) {
    return guard([&] {
        need_live(m);
        if (!fname)
            throw std::invalid_argument("null file name");
        m->m.dump(fname);
    });
}
int shamb200_model_load_dump(shamb200_model *m, const char *fname) {
    return guard([&] {
        need_live(m);
        if (!fname)
            throw std::invalid_argument("null file name");
        m->m.load_dump(fname);
    });
}
int shamb200_model_phantom_dump(shamb200_model *m, const char *fname) {
    return guard([&] {
        need_live(m);
        if (!fname)
            throw std::invalid_argument("null file name");
        m->m.phantom_dump(fname);
    });
}
int shamb200_model_init_from_phantom_dump(shamb200_model *m, const char *fname, double hpart_fact_load, uint64_t *kept) {
    return guard([&] {
        need_live(m);
        if (!fname)
            throw std::invalid_argument("null file name");
        const uint64_t k = m->m.init_from_phantom_dump(fname, hpart_fact_load);
        if (kept)
            *kept = k;
    });
}
int shamb200_model_vtk_dump(shamb200_model *m, const char *fname, int add_patch_world_id) {
    return guard([&] {
        need_live(m);
        if (!fname)
            throw std::invalid_argument("null file name");
        m->m.vtk_dump(fname, add_patch_world_id != 0);
    });
}
int shamb200_phantom_gen_config(const char *fname, int bypass_error, shamb200_solver_config *cfg) {
    return guard([&] {
        if (!fname || !cfg)
            throw std::invalid_argument("null argument");
        sb::phantom_gen_config(fname, bypass_error != 0, *cfg);
    });
}
int shamb200_phantom_copy(const char *fname_in, const char *fname_out) {
    return guard([&] {
        if (!fname_in || !fname_out)
            throw std::invalid_argument("null file name");
        sb::phantom_copy(fname_in, fname_out);
    });
}
int shamb200_phantom_header_float(const char *fname, const char *key, double *out, int *found) {
    return guard([&] {
        if (!fname || !key || !out || !found)
            throw std::invalid_argument("null argument");
        int64_t dummy = 0;
        *found        = sb::phantom_header(fname, key, 0, out, &dummy);
    });
}
int shamb200_phantom_header_int(const char *fname, const char *key, int64_t *out, int *found) {
    return guard([&] {
        if (!fname || !key || !out || !found)
            throw std::invalid_argument("null argument");
        double dummy = 0;
        *found       = sb::phantom_header(fname, key, 1, &dummy, out);
    });
}
int shamb200_phantom_compare(const char *fname_a, const char *fname_b, uint64_t *offenses) {
    return guard([&] {
        if (!fname_a || !fname_b || !offenses)
            throw std::invalid_argument("null argument");
        *offenses = sb::phantom_compare(fname_a, fname_b);
    });
}
int shamb200_model_init_scheduler(shamb200_model *m, uint64_t crit_split, uint64_t crit_merge, uint32_t step_freq) {
    return guard([&] {
        need_live(m);
        m->m.crit_split     = crit_split;
        m->m.crit_merge     = crit_merge;
        m->m.scheduler_freq = step_freq;
    });
}
int shamb200_model_scheduler_step(shamb200_model *m, int do_split_merge, int do_load_balancing) {
    return guard([&] {
        need_live(m);
        m->m.scheduler_step(do_split_merge != 0, do_load_balancing != 0);
    });
}
int shamb200_model_split_patch(shamb200_model *m, uint32_t ip) {
    return guard([&] {
        need_live(m);
        m->m.split_patch(ip);
    });
}
int shamb200_model_merge_patches(shamb200_model *m, uint32_t ip0) {
    return guard([&] {
        need_live(m);
        m->m.merge_patches(ip0);
    });
}
int shamb200_model_migrate_patch(shamb200_model *m, uint32_t ip, int new_owner) {
    return guard([&] {
        need_live(m);
        m->m.refresh_counts();
        m->m.migrate_patch(ip, new_owner);
    });
}
int shamb200_model_patch_info(shamb200_model *m, uint32_t ip, uint64_t out[8], double box[6]) {
    return guard([&] {
        need_live(m);
        const PatchD &p = m->m.patches.at(ip);
        out[0] = p.id;
        for (int d = 0; d < 3; d++) {
            out[1 + d] = p.cmin[d];
            out[4 + d] = p.cmax[d];
            if (box) {
                box[d]     = p.lo[d];
                box[3 + d] = p.hi[d];
            }
        }
        out[7] = u64(p.owner);
    });
}
int shamb200_model_scheduler_log(shamb200_model *m, double out[8]) {
    return guard([&] {
        need_live(m);
        const auto &l = m->m.sched_log;
        out[0] = l.splits, out[1] = l.merges, out[2] = l.moves, out[3] = f64(l.moved_objects), out[4] = l.npatch;
        out[5] = f64(l.max_rank_load), out[6] = l.mean_rank_load;
        out[7] = l.mean_rank_load > 0 ? f64(l.max_rank_load) / l.mean_rank_load - 1. : 0.;
    });
}
int shamb200_model_add_lattice_hcp(shamb200_model *m, double dr, const double box_min[3], const double box_max[3], uint64_t *added) {
    return guard([&] {
        need_live(m);
        u64 n = m->m.add_lattice_hcp(dr, box_min, box_max);
        if (added)
            *added = n;
    });
}
int shamb200_model_add_disc_lattice(shamb200_model *m, double dr, double r_in, double r_out, double zcut, uint64_t *added) {
    return guard([&] {
        need_live(m);
        u64 n = m->m.add_disc_lattice(dr, r_in, r_out, zcut);
        if (added)
            *added = n;
    });
}
int shamb200_model_add_disc_mc(shamb200_model *m, uint64_t npart, uint64_t seed, double r_in, double r_out, double p,
                               double q, double H_r_in, double disc_mass, uint64_t *added) {
    return guard([&] {
        need_live(m);
        u64 n = m->m.add_disc_mc(npart, seed, r_in, r_out, p, q, H_r_in, disc_mass);
        if (added)
            *added = n;
    });
}
int shamb200_model_set_value_in_a_box(shamb200_model *m, const char *field, int ivar, double val,
                                      const double box_min[3], const double box_max[3]) {
    return guard([&] {
        need_live(m);
        m->m.set_value_in_a_box(field, ivar, val, box_min, box_max);
    });
}
int shamb200_model_set_value_in_sphere(shamb200_model *m, const char *field, double val, const double center[3], double radius) {
    return guard([&] {
        need_live(m);
        m->m.set_value_in_sphere(field, val, center, radius);
    });
}
int shamb200_model_add_kernel_value(shamb200_model *m, const char *field, double val, const double center[3], double h_ker) {
    return guard([&] {
        need_live(m);
        m->m.add_kernel_value(field, val, center, h_ker);
    });
}
int shamb200_model_get_sum(shamb200_model *m, const char *field, double out[3]) {
    return guard([&] {
        need_live(m);
        m->m.get_sum(field, out);
    });
}
int shamb200_model_total_part_count(shamb200_model *m, uint64_t *out) {
    return guard([&] {
        need_live(m);
        *out = m->m.total_part_count();
    });
}
int shamb200_model_set_particle_mass(shamb200_model *m, double gpart_mass) {
    return guard([&] {
        need_live(m);
        m->m.cfg.gpart_mass = gpart_mass;
    });
}
uint32_t shamb200_model_patch_count(shamb200_model *m) {
    return model_is_live(m) ? (uint32_t) m->m.patches.size() : 0;
}
int shamb200_model_patch_is_local(shamb200_model *m, uint32_t ip) {
    return model_is_live(m) && ip < m->m.patches.size() && m->m.is_local(m->m.patches[ip]);
}
uint32_t shamb200_model_patch_size(shamb200_model *m, uint32_t ip) {
    if (!model_is_live(m) || ip >= m->m.patches.size() || !m->m.is_local(m->m.patches[ip]))
        return 0;
    return m->m.patches[ip].f.n;
}
int64_t shamb200_model_get(shamb200_model *m, uint32_t ip, const char *name, void *out, int64_t cap_bytes) {
    int64_t r = -1;
    int rc    = guard([&] {
        need_live(m);
        r = m->m.get(ip, name, out, cap_bytes);
    });
    return rc == SHAMB200_OK ? r : -2;
}
int shamb200_model_set_field(shamb200_model *m, uint32_t ip, const char *name, const double *in, uint64_t count) {
    return guard([&] {
        need_live(m);
        m->m.set_field(ip, name, in, count);
    });
}
int shamb200_model_evolve_once(shamb200_model *m) {
    return guard([&] {
        need_live(m);
        m->m.evolve_once();
    });
}
int shamb200_model_evolve_once_host(shamb200_model *m, uint32_t ip, const shamb200_host_patchdata *in,
                                    shamb200_host_patchdata *out) {
    return guard([&] {
        need_live(m);
        m->m.evolve_once_host(ip, in, out);
    });
}
int shamb200_host_register(void *p, uint64_t bytes) {
    return guard([&] { SB_CUDA_CHECK(cudaHostRegister(p, size_t(bytes), cudaHostRegisterDefault)); });
}
int shamb200_host_unregister(void *p) {
    return guard([&] { SB_CUDA_CHECK(cudaHostUnregister(p)); });
}
int shamb200_model_search_stats(shamb200_model *m, uint64_t out[2]) {
    return guard([&] {
        need_live(m);
        out[0] = m->m.K_local;
        out[1] = m->m.pair_tests_local;
    });
}
int shamb200_model_host_traffic(shamb200_model *m, uint64_t out[2]) {
    return guard([&] {
        need_live(m);
        out[0] = m->m.pipe.bytes_h2d;
        out[1] = m->m.pipe.bytes_d2h;
    });
}
int shamb200_model_list_tolerance(shamb200_model *m, double out[4]) {
    return guard([&] {
        need_live(m);
        out[0] = m->m.list_tol_last;
        out[1] = m->m.h_growth_last;
        out[2] = m->m.list_tol_next > 1. ? m->m.list_tol_next : m->m.cfg.htol_up_coarse_cycle;
        out[3] = double(m->m.list_fallbacks);
    });
}
int shamb200_model_host_step_info(shamb200_model *m, uint32_t ip, uint64_t out[2]) {
    return guard([&] {
        need_live(m);
        out[0] = m->m.pipe.nslices;
        (void) m->m.patches.at(ip);
        out[1] = m->m.pipe.far_host.p ? m->m.pipe.far_host.p[0] : 0;
    });
}
int shamb200_model_state(shamb200_model *m, double out[12]) {
    return guard([&] {
        need_live(m);
        Model &M = m->m;
        u64 nloc = 0;
        for (auto &p : M.patches)
            if (M.is_local(p))
                nloc += p.f.n;
        out[0]  = M.time;
        out[1]  = M.dt;
        out[2]  = M.cfl_multiplier;
        out[3]  = M.eps_v;
        out[4]  = M.h_subcycles;
        out[5]  = M.h_iters_last;
        out[6]  = M.corrector_iter;
        out[7]  = f64(M.npart_all);
        out[8]  = M.t_step;
        out[9]  = M.t_step > 0 ? f64(nloc) / M.t_step : 0;
        out[10] = f64(M.K_local);
        out[11] = f64(nloc);
    });
}
int shamb200_model_conservation(shamb200_model *m, double out[8]) {
    return guard([&] {
        need_live(m);
        for (int k = 0; k < 8; k++)
            out[k] = m->m.conservation[k];
    });
}
int shamb200_model_set_next_dt(shamb200_model *m, double dt) {
    return guard([&] {
        need_live(m);
        m->m.dt = dt;
    });
}
int shamb200_model_set_time(shamb200_model *m, double t) {
    return guard([&] {
        need_live(m);
        m->m.time = t;
    });
}
int shamb200_model_reorder_particles(shamb200_model *m) {
    return guard([&] {
        need_live(m);
        m->m.reorder_particles();
        SB_CUDA_CHECK(cudaStreamSynchronize(m->m.s()));
    });
}
int shamb200_model_set_cfl_multiplier(shamb200_model *m, double v) {
    return guard([&] {
        need_live(m);
        m->m.cfl_multiplier = v;
    });
}
int shamb200_model_stage_times(shamb200_model *m, const char **names, const double **ms, uint32_t *count) {
    return guard([&] {
        need_live(m);
        *names = m->m.timer.names_joined.c_str();
        *ms    = m->m.timer.values.data();
        *count = (uint32_t) m->m.timer.values.size();
    });
}

// NCCL entry points live in solver_comm.cu
} // extern "C"
