// sph.cu — the neighbour-list loops of the SPH step (density/h Newton sweep, Ω, EOS, ∇·v, ∇×v,
// d(∇·v)/dt, AV switch, pressure+AV forces, v_sig/CFL) as hand-written CUDA for sm_100a.
//
// Strict build (-fmad=false, default): every expression is evaluated in the reference's order with
// separate IEEE multiplications and additions, so the results are bit-identical to the CPU oracle
// (oracle/sph_step.hpp, built with -ffp-contract=off).  The fast build (-DSB_FAST_MATH, FMA
// contraction allowed) is held to 1e-10 relative.
//
// Reference loops restated (paths relative to /root/reference/src/shammodels/sph):
//   src/modules/IterateSmoothingLengthDensity.cpp:52-119, src/modules/ComputeOmega.cpp:36-73,
//   src/modules/ComputeEos.cpp:147-248,54-130,724-800, src/modules/DiffOperator.cpp:85-128,200-260,
//   src/modules/DiffOperatorDtDivv.cpp:100-196,220-353, src/modules/UpdateViscosity.cpp:88-109,167-221,
//   src/modules/UpdateDerivs.cpp:176-271,667-763, src/modules/NodeUpdateDerivsVaryingAlphaAV.cpp:37-137,
//   include/shammodels/sph/math/{density,forces,q_ab}.hpp, src/Solver.cpp:2726-2788,
//   include/shammodels/sph/modules/ComputeCFL{Courant,Force}.hpp; kernels shammath/sphkernels.hpp.
//
// Work distribution: one thread per particle in sorted-Morton order (`order`), so the lanes of a warp
// gather overlapping neighbour records (32-byte Pack4 sectors) that stay L1/L2 resident.
#include "sph.cuh"
#include "sphkern.cuh"

namespace sb {

__device__ __forceinline__ Pack4 ldg4(const Pack4 *p) { return ld4(p); }

constexpr int SPH_BLOCK = 128;

/// block-level max/min of `v` (for valid lanes) folded into red[0] (max) / red[1] (min)
__device__ __forceinline__ void block_reduce_maxmin(f64 vmax, f64 vmin, u64 *red) {
    __shared__ f64 smax[SPH_BLOCK / 32], smin[SPH_BLOCK / 32];
    vmax = warp_max(vmax);
    vmin = warp_min(vmin);
    if ((threadIdx.x & 31) == 0) {
        smax[threadIdx.x >> 5] = vmax;
        smin[threadIdx.x >> 5] = vmin;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        f64 a = smax[0], b = smin[0];
#pragma unroll
        for (int k = 1; k < SPH_BLOCK / 32; k++) {
            a = fmax(a, smax[k]);
            b = fmin(b, smin[k]);
        }
        atomicMax((unsigned long long *) &red[0], (unsigned long long) f64_to_ordered(a));
        atomicMin((unsigned long long *) &red[1], (unsigned long long) f64_to_ordered(b));
    }
}

// =============================================================================================
// h Newton sweep (K20 + K21 fused: the eps max/min reduction happens in the same launch)
// =============================================================================================
template<class K>
__global__ void __launch_bounds__(SPH_BLOCK) h_iter_kernel(
    CsrView c, const f64 *__restrict__ xyz, size_t stride, const u32 *__restrict__ order, u32 n_order,
    const f64 *__restrict__ h_old, f64 *__restrict__ h_new, f64 *__restrict__ eps, f64 part_mass,
    f64 h_max_tot_max_evol, f64 h_max_evol_p, u64 *red) {
    using Kn = Kern<K>;
    u32 r    = blockIdx.x * blockDim.x + threadIdx.x;
    u32 id_a = 0xFFFFFFFFu;
    if (r < n_order)
        id_a = order ? order[r] : r;
    bool valid = id_a < c.N;
    f64 e_out  = 0;
    if (valid) {
        f64 e        = eps[id_a];
        e_out        = e;
        f64 h_max_evol_m = 1 / h_max_evol_p;
        if (e > 1e-6) {
            f64 ax = xyz[u64(id_a) * stride], ay = xyz[u64(id_a) * stride + 1], az = xyz[u64(id_a) * stride + 2];
            f64 h_a     = h_new[id_a];
            f64 dint    = h_a * h_a * Kn::Rkern * Kn::Rkern;
            f64 rho_sum = 0, sumdWdh = 0;
            u32 s0 = c.scanned[id_a], s1 = s0 + c.cnt[id_a];
            for (u32 k = s0; k < s1; k++) {
                u32 id_b = c.list[k];
                f64 dx = ax - xyz[u64(id_b) * stride], dy = ay - xyz[u64(id_b) * stride + 1],
                    dz = az - xyz[u64(id_b) * stride + 2];
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
            f64 ha_0 = h_old[id_a];
            if (new_h < ha_0 * h_max_tot_max_evol) {
                h_new[id_a] = new_h;
                e_out       = fabs(new_h - h_a) / ha_0;
            } else {
                h_new[id_a] = ha_0 * h_max_tot_max_evol;
                e_out       = -1;
            }
            eps[id_a] = e_out;
        }
    }
    block_reduce_maxmin(valid ? e_out : -f64(INFINITY), valid ? e_out : f64(INFINITY), red);
}

template<class K>
__global__ void __launch_bounds__(SPH_BLOCK) omega_kernel(
    CsrView c, const f64 *__restrict__ xyz, size_t stride, const u32 *__restrict__ order, u32 n_order,
    const f64 *__restrict__ hpart, f64 *__restrict__ omega, f64 part_mass) {
    using Kn = Kern<K>;
    u32 r    = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_order)
        return;
    u32 id_a = order ? order[r] : r;
    if (id_a >= c.N)
        return;
    f64 ax = xyz[u64(id_a) * stride], ay = xyz[u64(id_a) * stride + 1], az = xyz[u64(id_a) * stride + 2];
    f64 h_a     = hpart[id_a];
    f64 dint    = h_a * h_a * Kn::Rkern * Kn::Rkern;
    f64 rho_sum = 0, part_omega_sum = 0;
    u32 s0 = c.scanned[id_a], s1 = s0 + c.cnt[id_a];
    for (u32 k = s0; k < s1; k++) {
        u32 id_b = c.list[k];
        f64 dx = ax - xyz[u64(id_b) * stride], dy = ay - xyz[u64(id_b) * stride + 1],
            dz = az - xyz[u64(id_b) * stride + 2];
        f64 rab2 = dx * dx + dy * dy + dz * dz;
        if (rab2 > dint)
            continue;
        f64 rab = sqrt(rab2);
        rho_sum += part_mass * Kn::W_3d(rab, h_a);
        part_omega_sum += part_mass * Kn::dhW_3d(rab, h_a);
    }
    f64 rho_ha  = rho_h(part_mass, h_a, Kn::hfactd);
    omega[id_a] = 1 + (h_a / (3 * rho_ha)) * part_omega_sum;
}

// =============================================================================================
// EOS (K25) on the merged range: A = (x,y,z,h), B = (vx,vy,vz,u) → C.a = P, C.c = cs
// =============================================================================================
template<class K>
__global__ void __launch_bounds__(256) eos_kernel(
    int eos, const Pack4 *__restrict__ A, const Pack4 *__restrict__ B, Pack4 *__restrict__ C, u32 M,
    f64 pmass, f64 gamma, f64 cs0, f64 q, f64 r0) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M)
        return;
    Pack4 a = A[i];
    f64 rho = rho_h(pmass, a.d, K::hfactd);
    f64 P, cs;
    if (eos == EOSK_ADIABATIC) {
        f64 u = B[i].d;
        P     = (gamma - 1) * rho * u;
        cs    = sqrt(gamma * P / rho);
    } else if (eos == EOSK_ISOTHERMAL) {
        P  = cs0 * cs0 * rho;
        cs = cs0;
    } else {
        f64 r0sq  = r0 * r0;
        f64 mq    = -q;
        f64 Rsq   = a.a * a.a + a.b * a.b + a.c * a.c;
        f64 cs_sq = (cs0 * cs0) * pow(Rsq / r0sq, mq);
        cs        = sqrt(cs_sq);
        P         = cs_sq * rho;
    }
    C[i].a = P;
    C[i].c = cs;
}

// =============================================================================================
// ∇·v, ∇×v (K26)
// =============================================================================================
template<class K, bool CURL>
__global__ void __launch_bounds__(SPH_BLOCK) divcurl_kernel(
    CsrView c, const Pack4 *__restrict__ A, const Pack4 *__restrict__ B, const Pack4 *__restrict__ C,
    const u32 *__restrict__ order, u32 n_order, f64 pmass, f64 *__restrict__ divv, f64 *__restrict__ curlv) {
    using Kn = Kern<K>;
    u32 r    = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_order)
        return;
    u32 id_a = order ? order[r] : r;
    if (id_a >= c.N)
        return;
    constexpr f64 Rker2 = Kn::Rkern * Kn::Rkern;
    Pack4 pa = ldg4(A + id_a), va = ldg4(B + id_a);
    f64 h_a             = pa.d;
    f64 omega_a         = C[id_a].b;
    f64 rho_a           = rho_h(pmass, h_a, Kn::hfactd);
    f64 inv_rho_omega_a = 1. / (omega_a * rho_a);
    f64 lim_a           = h_a * h_a * Rker2;
    f64 sum_nabla_v     = 0;
    f64 cx = 0, cy = 0, cz = 0;
    u32 s0 = c.scanned[id_a], s1 = s0 + c.cnt[id_a];
    for (u32 k = s0; k < s1; k++) {
        u32 id_b = c.list[k];
        Pack4 pb = ldg4(A + id_b);
        f64 dx = pa.a - pb.a, dy = pa.b - pb.b, dz = pa.c - pb.c;
        f64 rab2 = dx * dx + dy * dy + dz * dz;
        f64 h_b  = pb.d;
        if (rab2 > lim_a && rab2 > h_b * h_b * Rker2)
            continue;
        f64 rab  = sqrt(rab2);
        Pack4 vb = ldg4(B + id_b);
        f64 vx = va.a - vb.a, vy = va.b - vb.b, vz = va.c - vb.c;
        f64 ux = dx / rab, uy = dy / rab, uz = dz / rab;
        if (rab < 1e-9) {
            ux = 0;
            uy = 0;
            uz = 0;
        }
        f64 dW = Kn::dW_3d(rab, h_a);
        f64 gx = dW * ux, gy = dW * uy, gz = dW * uz;
        sum_nabla_v += pmass * (vx * gx + vy * gy + vz * gz);
        if (CURL) {
            cx += pmass * (vy * gz - vz * gy);
            cy += pmass * (vz * gx - vx * gz);
            cz += pmass * (vx * gy - vy * gx);
        }
    }
    divv[id_a] = -inv_rho_omega_a * sum_nabla_v;
    if (CURL) {
        curlv[3 * u64(id_a)]     = -inv_rho_omega_a * cx;
        curlv[3 * u64(id_a) + 1] = -inv_rho_omega_a * cy;
        curlv[3 * u64(id_a) + 2] = -inv_rho_omega_a * cz;
    }
}

// =============================================================================================
// d(∇·v)/dt (K27), matrix form (Cullen & Dehnen 2010)
// =============================================================================================
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

template<class K, bool ALSO>
__global__ void __launch_bounds__(SPH_BLOCK) dtdivv_kernel(
    CsrView c, const Pack4 *__restrict__ A, const Pack4 *__restrict__ B, const Pack4 *__restrict__ D,
    const u32 *__restrict__ order, u32 n_order, f64 pmass, f64 *__restrict__ divv, f64 *__restrict__ curlv,
    f64 *__restrict__ dtdivv) {
    using Kn = Kern<K>;
    u32 r    = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_order)
        return;
    u32 id_a = order ? order[r] : r;
    if (id_a >= c.N)
        return;
    constexpr f64 Rker2 = Kn::Rkern * Kn::Rkern;
    Pack4 pa = ldg4(A + id_a), va = ldg4(B + id_a), aa = ldg4(D + id_a);
    f64 h_a   = pa.d;
    f64 lim_a = h_a * h_a * Rker2;
    M33 Rij, Rv, Ra;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            Rij.m[i][j] = 0;
            Rv.m[i][j]  = 0;
            Ra.m[i][j]  = 0;
        }
    u32 s0 = c.scanned[id_a], s1 = s0 + c.cnt[id_a];
    for (u32 k = s0; k < s1; k++) {
        u32 id_b = c.list[k];
        Pack4 pb = ldg4(A + id_b);
        f64 rx = pa.a - pb.a, ry = pa.b - pb.b, rz = pa.c - pb.c;
        f64 rab2 = rx * rx + ry * ry + rz * rz;
        f64 h_b  = pb.d;
        if (rab2 > lim_a && rab2 > h_b * h_b * Rker2)
            continue;
        f64 rab  = sqrt(rab2);
        Pack4 vb = ldg4(B + id_b), ab = ldg4(D + id_b);
        f64 v[3]  = {va.a - vb.a, va.b - vb.b, va.c - vb.c};
        f64 a[3]  = {aa.a - ab.a, aa.b - ab.b, aa.c - ab.c};
        f64 rr[3] = {rx, ry, rz};
        f64 ux = rx / rab, uy = ry / rab, uz = rz / rab;
        if (rab < 1e-9) {
            ux = 0;
            uy = 0;
            uz = 0;
        }
        f64 dW   = Kn::dW_3d(rab, h_a);
        f64 g[3] = {(dW * ux) * pmass, (dW * uy) * pmass, (dW * uz) * pmass}; // mdWab_b
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) {
                Rij.m[i][j] -= rr[i] * g[j]; // Rij_a[i] -= r_ab[i] * mdWab_b
                Rv.m[i][j] -= v[j] * g[i];   // Rij_a_dvk_dxj[i] -= v_ab * mdWab_b[i]
                Ra.m[i][j] -= a[j] * g[i];
            }
    }
    M33 inv = inv_33(Rij);
    M33 dv  = prod_33(inv, Rv);
    M33 da  = prod_33(inv, Ra);
    f64 div_ai = da.m[0][0] + da.m[1][1] + da.m[2][2];
    f64 tens   = dv.m[0][0] * dv.m[0][0] + dv.m[1][0] * dv.m[0][1] + dv.m[2][0] * dv.m[0][2]
               + dv.m[0][1] * dv.m[1][0] + dv.m[1][1] * dv.m[1][1] + dv.m[2][1] * dv.m[1][2]
               + dv.m[0][2] * dv.m[2][0] + dv.m[1][2] * dv.m[2][1] + dv.m[2][2] * dv.m[2][2];
    if (ALSO) {
        divv[id_a]               = dv.m[0][0] + dv.m[1][1] + dv.m[2][2];
        curlv[3 * u64(id_a)]     = dv.m[1][2] - dv.m[2][1];
        curlv[3 * u64(id_a) + 1] = dv.m[2][0] - dv.m[0][2];
        curlv[3 * u64(id_a) + 2] = dv.m[0][1] - dv.m[1][0];
    }
    dtdivv[id_a] = div_ai - tens;
}

// =============================================================================================
// AV switch (K28)
// =============================================================================================
__global__ void __launch_bounds__(256) av_update_kernel(
    int av, u32 N, f64 dt, f64 sigma_decay, f64 alpha_min, f64 alpha_max, const f64 *__restrict__ divv,
    const f64 *__restrict__ curlv, const f64 *__restrict__ dtdivv, const f64 *__restrict__ cs,
    const f64 *__restrict__ h, const f64 *__restrict__ alpha, f64 *__restrict__ alpha_updated) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    f64 cs_a = cs[i], h_a = h[i], alpha_a = alpha[i], divv_a = divv[i];
    f64 vsig            = cs_a;
    f64 inv_tau_a       = vsig * sigma_decay / h_a;
    f64 fact_t          = dt * inv_tau_a;
    f64 euler_impl_fact = 1 / (1 + fact_t);
    if (av == AVK_MM97) {
        f64 source       = fmax(0., -divv_a);
        f64 new_alpha    = (alpha_a + source * dt + fact_t * alpha_min) * euler_impl_fact;
        alpha_updated[i] = fmin(alpha_max, new_alpha);
    } else {
        const f64 eps_d = 2.220446049250313e-16;
        f64 cx = curlv[3 * u64(i)], cy = curlv[3 * u64(i) + 1], cz = curlv[3 * u64(i) + 2];
        f64 dtdivv_a = dtdivv[i];
        f64 fac      = fmax(-divv_a, 0.);
        fac *= fac;
        f64 traceS        = cx * cx + cy * cy + cz * cz;
        f64 balsara_corec = (fac + traceS > eps_d) ? fac / (fac + traceS) : 1.;
        f64 A_a           = balsara_corec * fmax(-dtdivv_a, 0.);
        f64 temp          = cs_a * cs_a;
        f64 alpha_loc_a   = fmin((cs_a > 0) ? 10 * h_a * h_a * A_a / (temp) : alpha_min, alpha_max);
        alpha_loc_a       = (temp > eps_d) ? alpha_loc_a : alpha_min;
        f64 new_alpha     = (alpha_a + alpha_loc_a * fact_t) * euler_impl_fact;
        if (alpha_loc_a > alpha_a)
            new_alpha = alpha_loc_a;
        alpha_updated[i] = new_alpha;
    }
}

// =============================================================================================
// Forces (K29 + K30's `axyz += axyz_ext`)
//   A = (x,y,z,h)  B = (vx,vy,vz,u)  C = (P, omega, cs, alpha)
// =============================================================================================
template<class K, int AV>
__global__ void __launch_bounds__(SPH_BLOCK) force_kernel(
    CsrView c, const Pack4 *__restrict__ A, const Pack4 *__restrict__ B, const Pack4 *__restrict__ C,
    const u32 *__restrict__ order, u32 n_order, SphParams p, const f64 *__restrict__ axyz_ext,
    f64 *__restrict__ axyz, f64 *__restrict__ duint) {
    using Kn = Kern<K>;
    u32 r    = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_order)
        return;
    u32 id_a = order ? order[r] : r;
    if (id_a >= c.N)
        return;
    constexpr f64 Rker2   = Kn::Rkern * Kn::Rkern;
    constexpr bool VARY   = (AV == AVK_MM97 || AV == AVK_CD10);
    constexpr bool DISC   = (AV == AVK_DISC);
    const f64 pmass       = p.pmass;
    Pack4 pa = ldg4(A + id_a), va = ldg4(B + id_a), ca = ldg4(C + id_a);
    f64 h_a = pa.d, u_a = va.d, P_a = ca.a, omega_a = ca.b, cs_a = ca.c;
    f64 alpha_a           = VARY ? ca.d : p.alpha_AV;
    f64 rho_a             = rho_h(pmass, h_a, Kn::hfactd);
    f64 rho_a_sq          = rho_a * rho_a;
    f64 rho_a_inv         = 1. / rho_a;
    f64 omega_a_rho_a_inv = 1 / (omega_a * rho_a);
    f64 lim_a             = h_a * h_a * Rker2;
    f64 fx = 0, fy = 0, fz = 0, dU = 0;
    u32 s0 = c.scanned[id_a], s1 = s0 + c.cnt[id_a];
    for (u32 k = s0; k < s1; k++) {
        u32 id_b = c.list[k];
        Pack4 pb = ldg4(A + id_b);
        f64 dx = pa.a - pb.a, dy = pa.b - pb.b, dz = pa.c - pb.c;
        f64 rab2 = dx * dx + dy * dy + dz * dz;
        f64 h_b  = pb.d;
        if (rab2 > lim_a && rab2 > h_b * h_b * Rker2)
            continue;
        f64 rab  = sqrt(rab2);
        Pack4 vb = ldg4(B + id_b), cb = ldg4(C + id_b);
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
        // vsig_u (forces.hpp:38-45)
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
        // add_to_derivs_sph_artif_visco_cond (forces.hpp:171-224)
        f64 AV_P_a = P_a + qa_ab;
        f64 AV_P_b = P_b + qb_ab;
        f64 rho_b_sq   = rho_b * rho_b;
        f64 sub_fact_a = rho_a_sq * omega_a;
        f64 sub_fact_b = rho_b_sq * omega_b;
        f64 ka = (AV_P_a) *inv_sat_zero(sub_fact_a);
        f64 kb = (AV_P_b) *inv_sat_zero(sub_fact_b);
        f64 gax = ux * Fab_a, gay = uy * Fab_a, gaz = uz * Fab_a; // r_ab_unit * Fab_a
        f64 gbx = ux * Fab_b, gby = uy * Fab_b, gbz = uz * Fab_b;
        fx += -pmass * (ka * gax + kb * gbx);
        fy += -pmass * (ka * gay + kb * gby);
        fz += -pmass * (ka * gaz + kb * gbz);
        dU += AV_P_a * (omega_a_rho_a_inv * rho_a_inv) * pmass * (vx * gax + vy * gay + vz * gaz);
        dU += pmass * p.alpha_u * vsig_u * (u_a - u_b) * 0.5
              * (Fab_a * omega_a_rho_a_inv + Fab_b / (rho_b * omega_b));
    }
    axyz[3 * u64(id_a)]     = fx + axyz_ext[3 * u64(id_a)];
    axyz[3 * u64(id_a) + 1] = fy + axyz_ext[3 * u64(id_a) + 1];
    axyz[3 * u64(id_a) + 2] = fz + axyz_ext[3 * u64(id_a) + 2];
    duint[id_a]             = dU;
}

// =============================================================================================
// v_sig (K33) + CFL (K34) + min reduction
// =============================================================================================
template<class K>
__global__ void __launch_bounds__(SPH_BLOCK) vsig_cfl_kernel(
    CsrView c, const Pack4 *__restrict__ A, const Pack4 *__restrict__ B, const Pack4 *__restrict__ C,
    const u32 *__restrict__ order, u32 n_order, const f64 *__restrict__ axyz, f64 C_cour, f64 C_force,
    f64 *__restrict__ vsig_out, f64 *__restrict__ cfl_out, u64 *red_min) {
    using Kn = Kern<K>;
    u32 r    = blockIdx.x * blockDim.x + threadIdx.x;
    u32 id_a = 0xFFFFFFFFu;
    if (r < n_order)
        id_a = order ? order[r] : r;
    bool valid = id_a < c.N;
    f64 dt_out = f64(INFINITY);
    if (valid) {
        constexpr f64 Rker2 = Kn::Rkern * Kn::Rkern;
        Pack4 pa = ldg4(A + id_a), va = ldg4(B + id_a);
        f64 h_a = pa.d, cs_a = C[id_a].c;
        f64 lim_a    = h_a * h_a * Rker2;
        f64 vsig_max = 0;
        u32 s0 = c.scanned[id_a], s1 = s0 + c.cnt[id_a];
        for (u32 k = s0; k < s1; k++) {
            u32 id_b = c.list[k];
            Pack4 pb = ldg4(A + id_b);
            f64 dx = pa.a - pb.a, dy = pa.b - pb.b, dz = pa.c - pb.c;
            f64 rab2 = dx * dx + dy * dy + dz * dz;
            f64 h_b  = pb.d;
            if (rab2 > lim_a && rab2 > h_b * h_b * Rker2)
                continue;
            f64 rab  = sqrt(rab2);
            Pack4 vb = ldg4(B + id_b);
            f64 vx = va.a - vb.a, vy = va.b - vb.b, vz = va.c - vb.c;
            f64 ux = dx / rab, uy = dy / rab, uz = dz / rab;
            if (rab < 1e-9) {
                ux = 0;
                uy = 0;
                uz = 0;
            }
            f64 abs_v_ab_r_ab = fabs(vx * ux + vy * uy + vz * uz);
            f64 vsig_a        = 1.0 * cs_a + 2.0 * abs_v_ab_r_ab;
            vsig_max          = fmax(vsig_max, vsig_a);
        }
        vsig_out[id_a] = vsig_max;
        f64 dt_c = C_cour * h_a / vsig_max;
        f64 ax = axyz[3 * u64(id_a)], ay = axyz[3 * u64(id_a) + 1], az = axyz[3 * u64(id_a) + 2];
        f64 dt_f = C_force * sqrt(h_a / sqrt(ax * ax + ay * ay + az * az));
        dt_out   = fmin(fmin(f64(INFINITY), dt_c), dt_f);
        cfl_out[id_a] = dt_out;
    }
    __shared__ f64 smin[SPH_BLOCK / 32];
    f64 m = warp_min(dt_out);
    if ((threadIdx.x & 31) == 0)
        smin[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        f64 b = smin[0];
#pragma unroll
        for (int k = 1; k < SPH_BLOCK / 32; k++)
            b = fmin(b, smin[k]);
        atomicMin((unsigned long long *) red_min, (unsigned long long) f64_to_ordered(b));
    }
}

// =============================================================================================
// launchers
// =============================================================================================
#define SB_KDISPATCH(kernel, CALL)                                                               \
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

void h_iterate(
    cudaStream_t s, int kernel, CsrView c, const f64 *xyz, size_t stride, const u32 *order, u32 n_order,
    const f64 *h_old, f64 *h_new, f64 *eps, f64 pmass, f64 h_evol_max, f64 h_evol_iter_max, u64 *red) {
    if (!order)
        n_order = c.N;
    if (!n_order)
        return;
    SB_KDISPATCH(kernel, (h_iter_kernel<KT><<<grid_for(n_order, SPH_BLOCK), SPH_BLOCK, 0, s>>>(
                             c, xyz, stride, order, n_order, h_old, h_new, eps, pmass, h_evol_max,
                             h_evol_iter_max, red)));
}

void compute_omega(
    cudaStream_t s, int kernel, CsrView c, const f64 *xyz, size_t stride, const u32 *order, u32 n_order,
    const f64 *hpart, f64 *omega, f64 pmass) {
    if (!order)
        n_order = c.N;
    if (!n_order)
        return;
    SB_KDISPATCH(kernel, (omega_kernel<KT><<<grid_for(n_order, SPH_BLOCK), SPH_BLOCK, 0, s>>>(
                             c, xyz, stride, order, n_order, hpart, omega, pmass)));
}

void compute_eos(
    cudaStream_t s, int kernel, int eos, const Pack4 *A, const Pack4 *B, Pack4 *C, u32 M, f64 pmass,
    f64 gamma, f64 cs0, f64 q, f64 r0) {
    if (!M)
        return;
    SB_KDISPATCH(kernel, (eos_kernel<KT><<<grid_for(M, 256), 256, 0, s>>>(eos, A, B, C, M, pmass, gamma, cs0, q, r0)));
}

void compute_divv_curlv(
    cudaStream_t s, int kernel, CsrView c, const Pack4 *A, const Pack4 *B, const Pack4 *C, const u32 *order,
    u32 n_order, f64 pmass, f64 *divv, f64 *curlv) {
    if (!order)
        n_order = c.N;
    if (!n_order)
        return;
    if (curlv)
        SB_KDISPATCH(kernel, (divcurl_kernel<KT, true><<<grid_for(n_order, SPH_BLOCK), SPH_BLOCK, 0, s>>>(
                                 c, A, B, C, order, n_order, pmass, divv, curlv)));
    else
        SB_KDISPATCH(kernel, (divcurl_kernel<KT, false><<<grid_for(n_order, SPH_BLOCK), SPH_BLOCK, 0, s>>>(
                                 c, A, B, C, order, n_order, pmass, divv, curlv)));
}

void compute_dtdivv(
    cudaStream_t s, int kernel, CsrView c, const Pack4 *A, const Pack4 *B, const Pack4 *D, const u32 *order,
    u32 n_order, f64 pmass, bool also, f64 *divv, f64 *curlv, f64 *dtdivv) {
    if (!order)
        n_order = c.N;
    if (!n_order)
        return;
    if (also)
        SB_KDISPATCH(kernel, (dtdivv_kernel<KT, true><<<grid_for(n_order, SPH_BLOCK), SPH_BLOCK, 0, s>>>(
                                 c, A, B, D, order, n_order, pmass, divv, curlv, dtdivv)));
    else
        SB_KDISPATCH(kernel, (dtdivv_kernel<KT, false><<<grid_for(n_order, SPH_BLOCK), SPH_BLOCK, 0, s>>>(
                                 c, A, B, D, order, n_order, pmass, divv, curlv, dtdivv)));
}

void update_av(
    cudaStream_t s, int av, u32 N, f64 dt, f64 sigma_decay, f64 alpha_min, f64 alpha_max, const f64 *divv,
    const f64 *curlv, const f64 *dtdivv, const f64 *cs, const f64 *h, const f64 *alpha, f64 *alpha_updated) {
    if (!N)
        return;
    av_update_kernel<<<grid_for(N, 256), 256, 0, s>>>(
        av, N, dt, sigma_decay, alpha_min, alpha_max, divv, curlv, dtdivv, cs, h, alpha, alpha_updated);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

void compute_forces(
    cudaStream_t s, int kernel, int av, CsrView c, const Pack4 *A, const Pack4 *B, const Pack4 *C,
    const u32 *order, u32 n_order, SphParams p, const f64 *axyz_ext, f64 *axyz, f64 *duint) {
    if (!order)
        n_order = c.N;
    if (!n_order)
        return;
    unsigned g = grid_for(n_order, SPH_BLOCK);
    switch (av) {
    case AVK_CONSTANT:
        SB_KDISPATCH(kernel, (force_kernel<KT, AVK_CONSTANT><<<g, SPH_BLOCK, 0, s>>>(c, A, B, C, order, n_order, p, axyz_ext, axyz, duint)));
        break;
    case AVK_MM97:
    case AVK_CD10:
        SB_KDISPATCH(kernel, (force_kernel<KT, AVK_CD10><<<g, SPH_BLOCK, 0, s>>>(c, A, B, C, order, n_order, p, axyz_ext, axyz, duint)));
        break;
    case AVK_DISC:
        SB_KDISPATCH(kernel, (force_kernel<KT, AVK_DISC><<<g, SPH_BLOCK, 0, s>>>(c, A, B, C, order, n_order, p, axyz_ext, axyz, duint)));
        break;
    default: throw std::invalid_argument("unsupported artificial viscosity configuration");
    }
}

void compute_vsig_cfl(
    cudaStream_t s, int kernel, CsrView c, const Pack4 *A, const Pack4 *B, const Pack4 *C, const u32 *order,
    u32 n_order, const f64 *axyz, f64 C_cour, f64 C_force, f64 *vsig, f64 *cfl_dt, u64 *red_min) {
    if (!order)
        n_order = c.N;
    if (!n_order)
        return;
    SB_KDISPATCH(kernel, (vsig_cfl_kernel<KT><<<grid_for(n_order, SPH_BLOCK), SPH_BLOCK, 0, s>>>(
                             c, A, B, C, order, n_order, axyz, C_cour, C_force, vsig, cfl_dt, red_min)));
}

} // namespace sb
