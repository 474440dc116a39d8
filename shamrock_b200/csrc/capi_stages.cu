// capi_stages.cu — stage-level entry points of the SPH modules (include/shamb200.h, "SPH modules on
// merged patch data"): EOS, div / curl v, d(div v)/dt, AV switch, forces, v_sig / CFL, leapfrog.
// Each packs the caller's merged fields into the 32-byte records of sph.cu (A = x y z h, B = v u,
// C = P Ω c_s α, D = a) in a context arena and runs the id-ordered loops of sph.cu (-fmad=false: the
// reference's expressions in the reference's order).
#include "solver.cuh"
#include "stream_kernels.cuh"
#include <cstring>
#include <limits>

using namespace sb;

namespace {
template<class F>
int guard(F &&f) {
    try {
        f();
        return SHAMB200_OK;
    } catch (const CudaError &e) {
        set_last_error(e.what());
        return SHAMB200_ERR_CUDA;
    } catch (const std::invalid_argument &e) {
        set_last_error(e.what());
        return SHAMB200_ERR_INVALID;
    } catch (const std::exception &e) {
        set_last_error(e.what());
        return SHAMB200_ERR_RUNTIME;
    }
}
f64 bits_to_f64(u64 b) {
    f64 d;
    memcpy(&d, &b, 8);
    return d;
}

/// merged fields -> records; a NULL field leaves zeros (Ω: 1)
__global__ void __launch_bounds__(256) pack_merged_kernel(
    u32 M, const f64 *__restrict__ xyz, size_t stride, const f64 *__restrict__ h, const f64 *__restrict__ vxyz,
    const f64 *__restrict__ u, const f64 *__restrict__ axyz, const f64 *__restrict__ omega,
    const f64 *__restrict__ P, const f64 *__restrict__ cs, const f64 *__restrict__ alpha, Pack4 *__restrict__ A,
    Pack4 *__restrict__ B, Pack4 *__restrict__ C, Pack4 *__restrict__ D) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M)
        return;
    A[i] = Pack4{xyz[i * stride], xyz[i * stride + 1], xyz[i * stride + 2], h[i]};
    B[i] = Pack4{vxyz ? vxyz[3 * u64(i)] : 0., vxyz ? vxyz[3 * u64(i) + 1] : 0., vxyz ? vxyz[3 * u64(i) + 2] : 0.,
                 u ? u[i] : 0.};
    C[i] = Pack4{P ? P[i] : 0., omega ? omega[i] : 1., cs ? cs[i] : 0., alpha ? alpha[i] : 0.};
    if (D)
        D[i] = Pack4{axyz[3 * u64(i)], axyz[3 * u64(i) + 1], axyz[3 * u64(i) + 2], 0.};
}

void pack_merged(Ctx &c, const shamb200_merged_fields *f, bool want_D) {
    if (!f || !f->d_xyz || !f->d_hpart)
        throw std::invalid_argument("merged fields: xyz and hpart are required");
    if (f->stride_dbl != 3 && f->stride_dbl != 4)
        throw std::invalid_argument("merged fields: stride_dbl must be 3 or 4");
    if (f->real_cnt > f->obj_cnt)
        throw std::invalid_argument("merged fields: real_cnt exceeds obj_cnt");
    if (want_D && !f->d_axyz)
        throw std::invalid_argument("merged fields: this module reads axyz");
    SB_CUDA_CHECK(cudaSetDevice(c.device));
    const u32 M = f->obj_cnt;
    c.api_A.ensure(M, 1.1);
    c.api_B.ensure(M, 1.1);
    c.api_C.ensure(M, 1.1);
    if (want_D)
        c.api_D.ensure(M, 1.1);
    if (!M)
        return;
    pack_merged_kernel<<<grid_for(M, 256), 256, 0, c.stream>>>(
        M, f->d_xyz, f->stride_dbl, f->d_hpart, f->d_vxyz, f->d_uint, want_D ? f->d_axyz : nullptr, f->d_omega,
        f->d_pressure, f->d_soundspeed, f->d_alpha_AV, c.api_A.p, c.api_B.p, c.api_C.p, want_D ? c.api_D.p : nullptr);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

CsrView view_of(const shamb200_csr *csr, const shamb200_merged_fields *f) {
    if (!csr)
        throw std::invalid_argument("the neighbour cache is required");
    if (csr->obj_cnt != f->real_cnt)
        throw std::invalid_argument("the neighbour cache and the merged fields disagree on the object count");
    return CsrView{csr->d_cnt_neigh, csr->d_scanned_cnt, csr->d_index_neigh_map, csr->obj_cnt};
}

void need(const void *p, const char *what) {
    if (!p)
        throw std::invalid_argument(std::string("missing argument: ") + what);
}
int av_kind(int av) {
    switch (av) {
    case SHAMB200_AV_CONSTANT: return AVK_CONSTANT;
    case SHAMB200_AV_MM97: return AVK_MM97;
    case SHAMB200_AV_CD10: return AVK_CD10;
    case SHAMB200_AV_CONSTANT_DISC: return AVK_DISC;
    default: throw std::invalid_argument("unsupported artificial viscosity configuration");
    }
}
} // namespace

extern "C" {

int shamb200_compute_eos(
    shamb200_ctx *ctx, int kernel, int eos, const shamb200_merged_fields *f, double gpart_mass, double gamma,
    double cs0, double eos_q, double eos_r0, double *d_pressure, double *d_soundspeed) {
    return guard([&] {
        require_live(ctx);
        need(d_pressure, "d_pressure");
        need(d_soundspeed, "d_soundspeed");
        if (eos == SHAMB200_EOS_ADIABATIC)
            need(f ? f->d_uint : nullptr, "uint (adiabatic equation of state)");
        Ctx &c = ctx->c;
        pack_merged(c, f, false);
        compute_eos(c.stream, kernel, eos, c.api_A.p, c.api_B.p, c.api_C.p, f->obj_cnt, gpart_mass, gamma, cs0, eos_q, eos_r0);
        unpack_comp(c.stream, f->obj_cnt, c.api_C.p, 0, 1, d_pressure);
        unpack_comp(c.stream, f->obj_cnt, c.api_C.p, 2, 1, d_soundspeed);
    });
}

int shamb200_update_divv_curlv(
    shamb200_ctx *ctx, int kernel, const shamb200_csr *csr, const shamb200_merged_fields *f, double gpart_mass,
    double *d_divv, double *d_curlv) {
    return guard([&] {
        require_live(ctx);
        need(d_divv, "d_divv");
        need(f ? f->d_vxyz : nullptr, "vxyz");
        need(f->d_omega, "omega");
        Ctx &c = ctx->c;
        pack_merged(c, f, false);
        compute_divv_curlv(c.stream, kernel, view_of(csr, f), c.api_A.p, c.api_B.p, c.api_C.p, nullptr, 0, gpart_mass, d_divv, d_curlv);
    });
}

int shamb200_update_dtdivv(
    shamb200_ctx *ctx, int kernel, const shamb200_csr *csr, const shamb200_merged_fields *f, double gpart_mass,
    int also_divv_curlv, double *d_divv, double *d_curlv, double *d_dtdivv) {
    return guard([&] {
        require_live(ctx);
        need(d_dtdivv, "d_dtdivv");
        need(f ? f->d_vxyz : nullptr, "vxyz");
        if (also_divv_curlv) {
            need(d_divv, "d_divv");
            need(d_curlv, "d_curlv");
        }
        Ctx &c = ctx->c;
        pack_merged(c, f, true);
        compute_dtdivv(
            c.stream, kernel, view_of(csr, f), c.api_A.p, c.api_B.p, c.api_D.p, nullptr, 0, gpart_mass, also_divv_curlv != 0,
            d_divv, d_curlv, d_dtdivv);
    });
}

int shamb200_update_viscosity(
    shamb200_ctx *ctx, int av, uint32_t real_cnt, double dt, double sigma_decay, double alpha_min, double alpha_max,
    const double *d_divv, const double *d_curlv, const double *d_dtdivv, const double *d_soundspeed,
    const double *d_hpart, const double *d_alpha_AV, double *d_alpha_AV_updated) {
    return guard([&] {
        require_live(ctx);
        const int k = av_kind(av);
        if (k != AVK_MM97 && k != AVK_CD10)
            throw std::invalid_argument("the viscosity switch exists for MM97 and CD10 only");
        need(d_divv, "d_divv");
        need(d_soundspeed, "d_soundspeed");
        need(d_hpart, "d_hpart");
        need(d_alpha_AV, "d_alpha_AV");
        need(d_alpha_AV_updated, "d_alpha_AV_updated");
        if (k == AVK_CD10) {
            need(d_curlv, "d_curlv");
            need(d_dtdivv, "d_dtdivv");
        }
        SB_CUDA_CHECK(cudaSetDevice(ctx->c.device));
        update_av(
            ctx->c.stream, k, real_cnt, dt, sigma_decay, alpha_min, alpha_max, d_divv, d_curlv, d_dtdivv, d_soundspeed,
            d_hpart, d_alpha_AV, d_alpha_AV_updated);
    });
}

int shamb200_update_derivs(
    shamb200_ctx *ctx, int kernel, int av, const shamb200_csr *csr, const shamb200_merged_fields *f,
    double gpart_mass, double alpha_u, double alpha_AV, double beta_AV, const double *d_axyz_ext, double *d_axyz,
    double *d_duint) {
    return guard([&] {
        require_live(ctx);
        const int k = av_kind(av);
        need(d_axyz, "d_axyz");
        need(d_duint, "d_duint");
        need(f ? f->d_vxyz : nullptr, "vxyz");
        need(f->d_uint, "uint");
        need(f->d_omega, "omega");
        need(f->d_pressure, "pressure");
        need(f->d_soundspeed, "soundspeed");
        if (k == AVK_MM97 || k == AVK_CD10)
            need(f->d_alpha_AV, "alpha_AV (MM97 / CD10)");
        Ctx &c = ctx->c;
        pack_merged(c, f, false);
        const f64 *ext = d_axyz_ext;
        if (!ext) { // no external force: zeros
            c.api_tmp.ensure(size_t(f->real_cnt) * 3 + 1, 1.1);
            SB_CUDA_CHECK(cudaMemsetAsync(c.api_tmp.p, 0, size_t(f->real_cnt) * 3 * sizeof(f64), c.stream));
            ext = c.api_tmp.p;
        }
        SphParams sp{gpart_mass, alpha_u, alpha_AV, beta_AV};
        compute_forces(c.stream, kernel, k, view_of(csr, f), c.api_A.p, c.api_B.p, c.api_C.p, nullptr, 0, sp, ext, d_axyz, d_duint);
    });
}

int shamb200_vsig_cfl(
    shamb200_ctx *ctx, int kernel, const shamb200_csr *csr, const shamb200_merged_fields *f, const double *d_axyz,
    double C_cour, double C_force, double *d_vsig, double *d_cfl_dt, double *dt_min) {
    return guard([&] {
        require_live(ctx);
        need(d_axyz, "d_axyz");
        need(d_vsig, "d_vsig");
        need(d_cfl_dt, "d_cfl_dt");
        need(f ? f->d_vxyz : nullptr, "vxyz");
        need(f->d_soundspeed, "soundspeed");
        Ctx &c = ctx->c;
        pack_merged(c, f, false);
        c.red.ensure(8);
        c.h_red.ensure(8);
        c.h_red.p[0] = 0xFFFFFFFFFFFFFFFFull;
        SB_CUDA_CHECK(cudaMemcpyAsync(c.red.p, c.h_red.p, sizeof(u64), cudaMemcpyHostToDevice, c.stream));
        compute_vsig_cfl(
            c.stream, kernel, view_of(csr, f), c.api_A.p, c.api_B.p, c.api_C.p, nullptr, 0, d_axyz, C_cour, C_force,
            d_vsig, d_cfl_dt, c.red.p);
        SB_CUDA_CHECK(cudaMemcpyAsync(c.h_red.p + 1, c.red.p, sizeof(u64), cudaMemcpyDeviceToHost, c.stream));
        SB_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        if (dt_min)
            *dt_min = f->real_cnt ? ordered_to_f64(c.h_red.p[1]) : std::numeric_limits<f64>::infinity();
    });
}

int shamb200_leapfrog_predict(
    shamb200_ctx *ctx, uint32_t n, double dt, double *d_xyz, double *d_vxyz, const double *d_axyz, double *d_uint,
    const double *d_duint) {
    return guard([&] {
        require_live(ctx);
        need(d_xyz, "d_xyz");
        need(d_vxyz, "d_vxyz");
        need(d_axyz, "d_axyz");
        need(d_uint, "d_uint");
        need(d_duint, "d_duint");
        SB_CUDA_CHECK(cudaSetDevice(ctx->c.device));
        leapfrog_predictor(ctx->c.stream, n, dt, d_xyz, d_vxyz, d_axyz, d_uint, d_duint);
    });
}

int shamb200_leapfrog_correct(
    shamb200_ctx *ctx, uint32_t n, double half_dt, double *d_vxyz, const double *d_axyz, const double *d_axyz_old,
    double *d_uint, const double *d_duint, const double *d_duint_old, double out2[2]) {
    return guard([&] {
        require_live(ctx);
        need(d_vxyz, "d_vxyz");
        need(d_axyz, "d_axyz");
        need(d_axyz_old, "d_axyz_old");
        need(d_uint, "d_uint");
        need(d_duint, "d_duint");
        need(d_duint_old, "d_duint_old");
        Ctx &c = ctx->c;
        SB_CUDA_CHECK(cudaSetDevice(c.device));
        c.red.ensure(8);
        c.h_red.ensure(8);
        c.h_red.p[0] = 0; // ordered encoding: below every double
        c.h_red.p[1] = 0; // +0.0
        SB_CUDA_CHECK(cudaMemcpyAsync(c.red.p, c.h_red.p, 2 * sizeof(u64), cudaMemcpyHostToDevice, c.stream));
        leapfrog_corrector(
            c.stream, n, half_dt, d_vxyz, d_axyz, d_axyz_old, d_uint, d_duint, d_duint_old, c.red.p,
            reinterpret_cast<f64 *>(c.red.p + 1));
        SB_CUDA_CHECK(cudaMemcpyAsync(c.h_red.p + 2, c.red.p, 2 * sizeof(u64), cudaMemcpyDeviceToHost, c.stream));
        SB_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        if (out2) {
            out2[0] = n ? ordered_to_f64(c.h_red.p[2]) : 0.;
            out2[1] = bits_to_f64(c.h_red.p[3]);
        }
    });
}

} // extern "C"
