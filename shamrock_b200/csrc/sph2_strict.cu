// sph2_strict.cu — strict-fp SPH loops of the model path (compiled with -fmad=false).
//
// One thread per real particle in Morton order; the neighbour list is walked sequentially in ascending
// rank, i.e. in the order of the reference's ObjectCacheIterator (TreeTraversal.hpp:487-512), and every
// expression is the reference's, evaluated left to right with separate IEEE operations, so all outputs
// are bit-identical to the CPU oracle.  Reference loops restated (relative to
// /root/reference/src/shammodels/sph): src/modules/IterateSmoothingLengthDensity.cpp:52-119,
// src/modules/LoopSmoothingLengthIter.cpp:29-84, src/modules/ComputeOmega.cpp:36-73,
// src/modules/DiffOperator.cpp:85-128,200-260, src/modules/DiffOperatorDtDivv.cpp:100-196,220-353,
// src/modules/UpdateDerivs.cpp:176-271,667-763, src/modules/NodeUpdateDerivsVaryingAlphaAV.cpp:37-137,
// include/shammodels/sph/math/{density,forces,q_ab}.hpp, src/Solver.cpp:2726-2788,
// include/shammodels/sph/modules/ComputeCFL{Courant,Force}.hpp.
#include "sph2.cuh"
#include "sphkern.cuh"

namespace sb {

namespace {

constexpr int BLK = 128;

// ---- h Newton iteration (all sweeps) + Ω ------------------------------------------------------------
template<class K>
__global__ void __launch_bounds__(BLK) h_solve_kernel(
    RankCsr c, const Pack4 *__restrict__ SA, const f64 *__restrict__ h_old, f64 *__restrict__ hpart,
    f64 *__restrict__ eps, f64 *__restrict__ omega, f64 part_mass, f64 h_max_tot_max_evol, f64 h_max_evol_p,
    u32 max_sweeps, bool do_iter, bool do_omega, u64 *red) {
    using Kn   = Kern<K>;
    u32 k0     = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = k0 < c.count;
    f64 e_out  = 0;
    u32 sweeps = 0;
    if (valid) {
        u32 k, r, id;
        csr_item(c, k0, k, r, id);
        Pack4 a = ld4(SA + r);
        f64 h_a = hpart[id];
        u32 s0 = c.off[k], s1 = s0 + c.cnt[k];
        if (do_iter) {
            f64 e            = eps[id];
            f64 ha_0         = h_old[id];
            f64 h_max_evol_m = 1 / h_max_evol_p;
            while (sweeps < max_sweeps && e > 1e-6) {
                f64 dint    = h_a * h_a * Kn::Rkern * Kn::Rkern;
                f64 rho_sum = 0, sumdWdh = 0;
                for (u32 j = s0; j < s1; j++) {
                    Pack4 b = ld4(SA + c.list[j]);
                    f64 dx = a.a - b.a, dy = a.b - b.b, dz = a.c - b.c;
                    f64 rab2 = dx * dx + dy * dy + dz * dz;
                    if (rab2 > dint)
                        continue;
                    f64 rab = sqrt(rab2);
                    rho_sum += part_mass * Kn::W_3d(rab, h_a);
                    sumdWdh += part_mass * Kn::dhW_3d(rab, h_a);
                }
                f64 rho_ha  = rho_h(part_mass, h_a, Kn::hfactd);
                f64 f_iter  = rho_sum - rho_ha;
                f64 df_iter = sumdWdh + 3 * rho_ha / h_a;
                f64 new_h   = h_a - f_iter / df_iter;
                if (new_h < h_a * h_max_evol_m)
                    new_h = h_max_evol_m * h_a;
                if (new_h > h_a * h_max_evol_p)
                    new_h = h_max_evol_p * h_a;
                if (new_h < ha_0 * h_max_tot_max_evol) {
                    e   = fabs(new_h - h_a) / ha_0;
                    h_a = new_h;
                } else {
                    h_a = ha_0 * h_max_tot_max_evol;
                    e   = -1;
                }
                sweeps++;
            }
            e_out     = e;
            eps[id]   = e;
            hpart[id] = h_a;
        }
        if (do_omega) {
            f64 dint    = h_a * h_a * Kn::Rkern * Kn::Rkern;
            f64 rho_sum = 0, part_omega_sum = 0;
            for (u32 j = s0; j < s1; j++) {
                Pack4 b = ld4(SA + c.list[j]);
                f64 dx = a.a - b.a, dy = a.b - b.b, dz = a.c - b.c;
                f64 rab2 = dx * dx + dy * dy + dz * dz;
                if (rab2 > dint)
                    continue;
                f64 rab = sqrt(rab2);
                rho_sum += part_mass * Kn::W_3d(rab, h_a);
                part_omega_sum += part_mass * Kn::dhW_3d(rab, h_a);
            }
            f64 rho_ha = rho_h(part_mass, h_a, Kn::hfactd);
            omega[id]  = 1 + (h_a / (3 * rho_ha)) * part_omega_sum;
        }
    }
    if (do_iter) {
        __shared__ f64 smax[BLK / 32], smin[BLK / 32];
        __shared__ u32 ssw[BLK / 32];
        f64 vmax = warp_max(valid ? e_out : -f64(INFINITY));
        f64 vmin = warp_min(valid ? e_out : f64(INFINITY));
        u32 sw   = sweeps;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            sw = max(sw, __shfl_xor_sync(0xffffffffu, sw, o));
        if ((threadIdx.x & 31) == 0) {
            smax[threadIdx.x >> 5] = vmax;
            smin[threadIdx.x >> 5] = vmin;
            ssw[threadIdx.x >> 5]  = sw;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            f64 x = smax[0], y = smin[0];
            u32 z = ssw[0];
#pragma unroll
            for (int q = 1; q < BLK / 32; q++) {
                x = fmax(x, smax[q]);
                y = fmin(y, smin[q]);
                z = max(z, ssw[q]);
            }
            atomicMax((unsigned long long *) &red[0], (unsigned long long) f64_to_ordered(x));
            atomicMin((unsigned long long *) &red[1], (unsigned long long) f64_to_ordered(y));
            atomicMax((unsigned long long *) &red[2], (unsigned long long) z);
        }
    }
}

// ---- ∇·v, ∇×v, d(∇·v)/dt in one loop ------------------------------------------------------------------
struct M33 {
    f64 m[3][3]; // m[row] = the reference's std::array<Tvec,3>[row] = (x, y, z)
};
__device__ __forceinline__ M33 inv_33(const M33 &A) {
    f64 a00 = A.m[0][0], a10 = A.m[1][0], a20 = A.m[2][0];
    f64 a01 = A.m[0][1], a11 = A.m[1][1], a21 = A.m[2][1];
    f64 a02 = A.m[0][2], a12 = A.m[1][2], a22 = A.m[2][2];
    f64 det = (-a02 * a11 * a20 + a01 * a12 * a20 + a02 * a10 * a21 - a00 * a12 * a21 - a01 * a10 * a22
               + a00 * a11 * a22);
    M33 R;
    R.m[0][0] = (-a12 * a21 + a11 * a22) / det;
    R.m[0][1] = (a02 * a21 - a01 * a22) / det;
    R.m[0][2] = (-a02 * a11 + a01 * a12) / det;
    R.m[1][0] = (a12 * a20 - a10 * a22) / det;
    R.m[1][1] = (-a02 * a20 + a00 * a22) / det;
    R.m[1][2] = (a02 * a10 - a00 * a12) / det;
    R.m[2][0] = (-a11 * a20 + a10 * a21) / det;
    R.m[2][1] = (a01 * a20 - a00 * a21) / det;
    R.m[2][2] = (-a01 * a10 + a00 * a11) / det;
    return R;
}
__device__ __forceinline__ M33 prod_33(const M33 &A, const M33 &B) {
    M33 R;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            R.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j];
    return R;
}

/// SPHDIV: the SPH estimators of DiffOperator.cpp (divv, and curlv when CURL); MAT: the matrix form of
/// DiffOperatorDtDivv.cpp (dtdivv; also divv/curlv when COMBINED)
template<class K, bool SPHDIV, bool CURL, bool MAT, bool COMBINED>
__global__ void __launch_bounds__(BLK) av_operators_kernel(
    RankCsr c, const Pack4 *__restrict__ SA, const Pack4 *__restrict__ SB, const Pack4 *__restrict__ SC,
    const Pack4 *__restrict__ SD, f64 pmass, f64 *__restrict__ divv, f64 *__restrict__ curlv,
    f64 *__restrict__ dtdivv) {
    using Kn = Kern<K>;
    u32 k0   = blockIdx.x * blockDim.x + threadIdx.x;
    if (k0 >= c.count)
        return;
    u32 k, r, id;
    csr_item(c, k0, k, r, id);
    constexpr f64 Rker2 = Kn::Rkern * Kn::Rkern;
    Pack4 pa = ld4(SA + r), va = ld4(SB + r);
    Pack4 aa = MAT ? ld4(SD + r) : Pack4{0, 0, 0, 0};
    f64 h_a   = pa.d;
    f64 lim_a = h_a * h_a * Rker2;
    f64 sum_nabla_v = 0, cx = 0, cy = 0, cz = 0;
    M33 Rij, Rv, Ra;
    if (MAT) {
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) {
                Rij.m[i][j] = 0;
                Rv.m[i][j]  = 0;
                Ra.m[i][j]  = 0;
            }
    }
    u32 s0 = c.off[k], s1 = s0 + c.cnt[k];
    for (u32 j = s0; j < s1; j++) {
        u32 rb   = c.list[j];
        Pack4 pb = ld4(SA + rb);
        f64 dx = pa.a - pb.a, dy = pa.b - pb.b, dz = pa.c - pb.c;
        f64 rab2 = dx * dx + dy * dy + dz * dz;
        f64 h_b  = pb.d;
        if (rab2 > lim_a && rab2 > h_b * h_b * Rker2)
            continue;
        f64 rab  = sqrt(rab2);
        Pack4 vb = ld4(SB + rb);
        f64 vx = va.a - vb.a, vy = va.b - vb.b, vz = va.c - vb.c;
        f64 ux = dx / rab, uy = dy / rab, uz = dz / rab;
        if (rab < 1e-9) {
            ux = 0;
            uy = 0;
            uz = 0;
        }
        f64 dW = Kn::dW_3d(rab, h_a);
        if (SPHDIV) {
            f64 gx = dW * ux, gy = dW * uy, gz = dW * uz;
            sum_nabla_v += pmass * (vx * gx + vy * gy + vz * gz);
            if (CURL) {
                cx += pmass * (vy * gz - vz * gy);
                cy += pmass * (vz * gx - vx * gz);
                cz += pmass * (vx * gy - vy * gx);
            }
        }
        if (MAT) {
            Pack4 ab  = ld4(SD + rb);
            f64 v[3]  = {vx, vy, vz};
            f64 a[3]  = {aa.a - ab.a, aa.b - ab.b, aa.c - ab.c};
            f64 rr[3] = {dx, dy, dz};
            f64 g[3]  = {(dW * ux) * pmass, (dW * uy) * pmass, (dW * uz) * pmass}; // mdWab_b
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    Rij.m[i][q] -= rr[i] * g[q];
                    Rv.m[i][q] -= v[q] * g[i];
                    Ra.m[i][q] -= a[q] * g[i];
                }
        }
    }
    if (SPHDIV) {
        f64 omega_a         = SC[r].b;
        f64 rho_a           = rho_h(pmass, h_a, Kn::hfactd);
        f64 inv_rho_omega_a = 1. / (omega_a * rho_a);
        divv[id]            = -inv_rho_omega_a * sum_nabla_v;
        if (CURL) {
            curlv[3 * u64(id)]     = -inv_rho_omega_a * cx;
            curlv[3 * u64(id) + 1] = -inv_rho_omega_a * cy;
            curlv[3 * u64(id) + 2] = -inv_rho_omega_a * cz;
        }
    }
    if (MAT) {
        M33 inv = inv_33(Rij);
        M33 dv  = prod_33(inv, Rv);
        M33 da  = prod_33(inv, Ra);
        f64 div_ai = da.m[0][0] + da.m[1][1] + da.m[2][2];
        f64 tens   = dv.m[0][0] * dv.m[0][0] + dv.m[1][0] * dv.m[0][1] + dv.m[2][0] * dv.m[0][2]
                   + dv.m[0][1] * dv.m[1][0] + dv.m[1][1] * dv.m[1][1] + dv.m[2][1] * dv.m[1][2]
                   + dv.m[0][2] * dv.m[2][0] + dv.m[1][2] * dv.m[2][1] + dv.m[2][2] * dv.m[2][2];
        if (COMBINED) {
            divv[id]               = dv.m[0][0] + dv.m[1][1] + dv.m[2][2];
            curlv[3 * u64(id)]     = dv.m[1][2] - dv.m[2][1];
            curlv[3 * u64(id) + 1] = dv.m[2][0] - dv.m[0][2];
            curlv[3 * u64(id) + 2] = dv.m[0][1] - dv.m[1][0];
        }
        dtdivv[id] = div_ai - tens;
    }
}

// ---- forces + v_sig + CFL -------------------------------------------------------------------------------
template<class K, int AV>
__global__ void __launch_bounds__(BLK) force_cfl_kernel(
    RankCsr c, const Pack4 *__restrict__ SA, const Pack4 *__restrict__ SB, const Pack4 *__restrict__ SC,
    SphParams p, const f64 *__restrict__ axyz_ext, f64 *__restrict__ axyz, f64 *__restrict__ duint, f64 C_cour,
    f64 C_force, f64 *__restrict__ vsig_out, f64 *__restrict__ cfl_out, u64 *red_min) {
    using Kn   = Kern<K>;
    u32 k0     = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = k0 < c.count;
    f64 dt_out = f64(INFINITY);
    if (valid) {
        u32 k, r, id;
        csr_item(c, k0, k, r, id);
        constexpr f64 Rker2 = Kn::Rkern * Kn::Rkern;
        constexpr bool VARY = (AV == AVK_MM97 || AV == AVK_CD10);
        constexpr bool DISC = (AV == AVK_DISC);
        const f64 pmass     = p.pmass;
        Pack4 pa = ld4(SA + r), va = ld4(SB + r), ca = ld4(SC + r);
        f64 h_a = pa.d, u_a = va.d, P_a = ca.a, omega_a = ca.b, cs_a = ca.c;
        f64 alpha_a           = VARY ? ca.d : p.alpha_AV;
        f64 rho_a             = rho_h(pmass, h_a, Kn::hfactd);
        f64 rho_a_sq          = rho_a * rho_a;
        f64 rho_a_inv         = 1. / rho_a;
        f64 omega_a_rho_a_inv = 1 / (omega_a * rho_a);
        f64 lim_a             = h_a * h_a * Rker2;
        f64 fx = 0, fy = 0, fz = 0, dU = 0, vsig_max = 0;
        u32 s0 = c.off[k], s1 = s0 + c.cnt[k];
        for (u32 j = s0; j < s1; j++) {
            u32 rb   = c.list[j];
            Pack4 pb = ld4(SA + rb);
            f64 dx = pa.a - pb.a, dy = pa.b - pb.b, dz = pa.c - pb.c;
            f64 rab2 = dx * dx + dy * dy + dz * dz;
            f64 h_b  = pb.d;
            if (rab2 > lim_a && rab2 > h_b * h_b * Rker2)
                continue;
            f64 rab  = sqrt(rab2);
            Pack4 vb = ld4(SB + rb), cb = ld4(SC + rb);
            f64 u_b = vb.d, P_b = cb.a, omega_b = cb.b, cs_b = cb.c;
            f64 alpha_b = VARY ? cb.d : p.alpha_AV;
            f64 rho_b   = rho_h(pmass, h_b, Kn::hfactd);
            f64 Fab_a   = Kn::dW_3d(rab, h_a);
            f64 Fab_b   = Kn::dW_3d(rab, h_b);
            f64 vx = va.a - vb.a, vy = va.b - vb.b, vz = va.c - vb.c;
            f64 irab = inv_sat_positive(rab);
            f64 ux = dx * irab, uy = dy * irab, uz = dz * irab;
            f64 v_ab_r_ab     = vx * ux + vy * uy + vz * uz;
            f64 abs_v_ab_r_ab = fabs(v_ab_r_ab);
            f64 vsig_a        = alpha_a * cs_a + p.beta_AV * abs_v_ab_r_ab;
            f64 vsig_b        = alpha_b * cs_b + p.beta_AV * abs_v_ab_r_ab;
            f64 rho_avg = (rho_a + rho_b) * 0.5;
            f64 abs_dp  = fabs(P_a - P_b);
            f64 vsig_u  = sqrt(abs_dp / rho_avg);
            f64 qa_ab, qb_ab;
            if (DISC) { // q_av_disc (q_ab.hpp:42-60)
                f64 rabinv    = inv_sat_positive(rab);
                f64 prefact_a = -0.5 * rho_a * fabs(rabinv) * h_a;
                f64 vd_a      = (v_ab_r_ab < 0.) ? vsig_a : (alpha_a * cs_a);
                qa_ab         = prefact_a * vd_a * v_ab_r_ab;
                f64 prefact_b = -0.5 * rho_b * fabs(rabinv) * h_b;
                f64 vd_b      = (v_ab_r_ab < 0.) ? vsig_b : (alpha_b * cs_b);
                qb_ab         = prefact_b * vd_b * v_ab_r_ab;
            } else { // q_av (q_ab.hpp:37-40)
                qa_ab = fmax(-0.5 * rho_a * vsig_a * v_ab_r_ab, 0.);
                qb_ab = fmax(-0.5 * rho_b * vsig_b * v_ab_r_ab, 0.);
            }
            f64 AV_P_a     = P_a + qa_ab;
            f64 AV_P_b     = P_b + qb_ab;
            f64 rho_b_sq   = rho_b * rho_b;
            f64 sub_fact_a = rho_a_sq * omega_a;
            f64 sub_fact_b = rho_b_sq * omega_b;
            f64 ka = (AV_P_a) *inv_sat_zero(sub_fact_a);
            f64 kb = (AV_P_b) *inv_sat_zero(sub_fact_b);
            f64 gax = ux * Fab_a, gay = uy * Fab_a, gaz = uz * Fab_a;
            f64 gbx = ux * Fab_b, gby = uy * Fab_b, gbz = uz * Fab_b;
            fx += -pmass * (ka * gax + kb * gbx);
            fy += -pmass * (ka * gay + kb * gby);
            fz += -pmass * (ka * gaz + kb * gbz);
            dU += AV_P_a * (omega_a_rho_a_inv * rho_a_inv) * pmass * (vx * gax + vy * gay + vz * gaz);
            dU += pmass * p.alpha_u * vsig_u * (u_a - u_b) * 0.5
                  * (Fab_a * omega_a_rho_a_inv + Fab_b / (rho_b * omega_b));
            // v_sig of the CFL condition (Solver.cpp:2726-2788): its own unit vector (dr / rab), α=1, β=2
            f64 wx = dx / rab, wy = dy / rab, wz = dz / rab;
            if (rab < 1e-9) {
                wx = 0;
                wy = 0;
                wz = 0;
            }
            f64 abs_w = fabs(vx * wx + vy * wy + vz * wz);
            vsig_max  = fmax(vsig_max, 1.0 * cs_a + 2.0 * abs_w);
        }
        f64 ax = fx + axyz_ext[3 * u64(id)], ay = fy + axyz_ext[3 * u64(id) + 1], az = fz + axyz_ext[3 * u64(id) + 2];
        axyz[3 * u64(id)]     = ax;
        axyz[3 * u64(id) + 1] = ay;
        axyz[3 * u64(id) + 2] = az;
        duint[id]             = dU;
        vsig_out[id]          = vsig_max;
        f64 dt_c = C_cour * h_a / vsig_max;
        f64 dt_f = C_force * sqrt(h_a / sqrt(ax * ax + ay * ay + az * az));
        dt_out   = fmin(fmin(f64(INFINITY), dt_c), dt_f);
        cfl_out[id] = dt_out;
    }
    __shared__ f64 smin[BLK / 32];
    f64 m = warp_min(dt_out);
    if ((threadIdx.x & 31) == 0)
        smin[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        f64 b = smin[0];
#pragma unroll
        for (int q = 1; q < BLK / 32; q++)
            b = fmin(b, smin[q]);
        atomicMin((unsigned long long *) red_min, (unsigned long long) f64_to_ordered(b));
    }
}

} // namespace

#define SB_KD(kernel, CALL)                                                                      \
    do {                                                                                         \
        if ((kernel) == KERN_M4) {                                                               \
            using KT = KM4;                                                                      \
            CALL;                                                                                \
        } else {                                                                                 \
            using KT = KM6;                                                                      \
            CALL;                                                                                \
        }                                                                                        \
        SB_COUNT_LAUNCH();                                                                       \
        SB_LAUNCH_CHECK();                                                                       \
    } while (0)

void h_solve_strict(
    cudaStream_t s, int kernel, RankCsr c, const Pack4 *SA, const f64 *h_old, f64 *hpart, f64 *eps, f64 *omega,
    f64 pmass, f64 h_evol_max, f64 h_evol_iter_max, u32 max_sweeps, bool do_iter, bool do_omega, u64 *red) {
    if (!c.N || !c.count)
        return;
    SB_KD(kernel, (h_solve_kernel<KT><<<grid_for(c.count, BLK), BLK, 0, s>>>(
                      c, SA, h_old, hpart, eps, omega, pmass, h_evol_max, h_evol_iter_max, max_sweeps, do_iter,
                      do_omega, red)));
}

void av_operators_strict(
    cudaStream_t s, int kernel, RankCsr c, const Pack4 *SA, const Pack4 *SB, const Pack4 *SC, const Pack4 *SD,
    f64 pmass, bool want_curl, bool want_dtdivv, bool combined, f64 *divv, f64 *curlv, f64 *dtdivv) {
    if (!c.N || !c.count)
        return;
    unsigned g = grid_for(c.count, BLK);
#define AVOP(S_, C_, M_, CB_)                                                                    \
    SB_KD(kernel, (av_operators_kernel<KT, S_, C_, M_, CB_><<<g, BLK, 0, s>>>(c, SA, SB, SC, SD, pmass, divv, curlv, dtdivv)))
    if (want_dtdivv) {
        if (combined)
            AVOP(false, false, true, true);
        else if (want_curl)
            AVOP(true, true, true, false);
        else
            AVOP(true, false, true, false);
    } else {
        if (want_curl)
            AVOP(true, true, false, false);
        else
            AVOP(true, false, false, false);
    }
#undef AVOP
}

void force_cfl_strict(
    cudaStream_t s, int kernel, int av, RankCsr c, const Pack4 *SA, const Pack4 *SB, const Pack4 *SC, SphParams p,
    const f64 *axyz_ext, f64 *axyz, f64 *duint, f64 C_cour, f64 C_force, f64 *vsig, f64 *cfl_dt, u64 *red_min) {
    if (!c.N || !c.count)
        return;
    unsigned g = grid_for(c.count, BLK);
#define FRC(AV_)                                                                                 \
    SB_KD(kernel, (force_cfl_kernel<KT, AV_><<<g, BLK, 0, s>>>(                                  \
                      c, SA, SB, SC, p, axyz_ext, axyz, duint, C_cour, C_force, vsig, cfl_dt, red_min)))
    switch (av) {
    case AVK_CONSTANT: FRC(AVK_CONSTANT); break;
    case AVK_MM97:
    case AVK_CD10: FRC(AVK_CD10); break;
    case AVK_DISC: FRC(AVK_DISC); break;
    default: throw std::invalid_argument("unsupported artificial viscosity configuration");
    }
#undef FRC
}

} // namespace sb
