// sph2_fast.cu — fast-fp SPH loops of the model path (compiled with -fmad=true).
//
// Same algorithm and the same neighbour lists as sph2_strict.cu; what changes is how the FP64 pipe — the
// binding roof of these loops on B200 — is used:
//  * G lanes per particle stride over its neighbour list and combine their partial sums with warp
//    shuffles.  Measured on B200 (profiles/): with Morton-sorted records a small G wins (G = 2 for M4):
//    adjacent lanes are adjacent particles whose j-th neighbours sit in the same few cache lines, and
//    the per-particle prologue / reduction is amortised over more trips;
//  * everything that depends on ONE particle only (1/h, norm/h^4, rho, 1/(rho^2 Ω), 1/(rho Ω), α c_s,
//    P/(rho^2 Ω)) is computed once per particle (derive_fast) instead of once per pair: the pair math
//    has one rsqrt, one reciprocal and one sqrt left (the reference's has ~12 divisions / 2 sqrt);
//  * kernel normalisations and the particle mass are factored out of the sums; FMA contraction is on.
// Results agree with the strict path to rounding (~1e-15 relative per pair); the parity tests hold this
// mode to 1e-10 relative per particle (north-star tolerance).  Reference loops: see sph2_strict.cu.
#include "sph2.cuh"
#include "sphkern.cuh"
#include <cstdlib>

namespace sb {

namespace {

constexpr int BLK = 128;

/// sum over the G lanes of a group, result broadcast from the group's first lane (identical in all lanes)
/// lanes of this thread's group (groups of one warp may sit in different loop trips: the shuffles name
/// only the group's own lanes)
template<int G>
__device__ __forceinline__ u32 group_mask() {
    if (G >= 32)
        return 0xffffffffu;
    if (G == 1)
        return 1u << (threadIdx.x & 31u);
    u32 lane = threadIdx.x & 31u;
    return ((1u << G) - 1u) << (lane & ~u32(G - 1));
}
template<int G>
__device__ __forceinline__ f64 group_sum(f64 v) {
    const u32 m = group_mask<G>();
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1)
        v += __shfl_down_sync(m, v, o, G);
    return __shfl_sync(m, v, 0, G);
}
template<int G>
__device__ __forceinline__ f64 group_max(f64 v) {
    const u32 m = group_mask<G>();
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1)
        v = fmax(v, __shfl_down_sync(m, v, o, G));
    return __shfl_sync(m, v, 0, G);
}

// ---- branch-free pair math ------------------------------------------------------------------------------
/// 1/sqrt(x) for a normal positive x: MUFU.RSQ64H seed (~2^-20) + one third-order step (error ~ e^3),
/// five FP64 instructions and no special-case branch (CUDA's rsqrt() adds a range check and a slow path)
__device__ __forceinline__ f64 fast_rsqrt(f64 x) {
    f64 y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    f64 e = fma(-x, y * y, 1.0);
    f64 p = fma(e, 0.375, 0.5);
    return fma(p, y * e, y);
}
/// The kernel shape functions without branches: the piecewise polynomials of sphkernels.hpp are sums of
/// clamped powers, M4: f = (2-q)+^3 / 4 - (1-q)+^3, M6: f = (3-q)+^5 - 6 (2-q)+^5 + 15 (1-q)+^5, and
/// 2 x+ = x + |x| costs one add (the |.| is an operand modifier).  Same values as KM4 / KM6::f, df up to
/// rounding (~1e-16 absolute); no divergence between lanes sitting in different pieces.
template<class K>
struct FastK;
template<>
struct FastK<KM4> {
    /// df for any q >= 0
    __device__ static __forceinline__ f64 df(f64 q) {
        f64 t1 = 2. - q, t2 = 1. - q;
        f64 u1 = t1 + fabs(t1), u2 = t2 + fabs(t2); // 2 x+
        return fma(-0.1875, u1 * u1, 0.75 * (u2 * u2));
    }
    /// f for q <= 2
    __device__ static __forceinline__ f64 f_in(f64 q) {
        f64 t1 = 2. - q, t2 = 1. - q;
        f64 u2 = t2 + fabs(t2);
        return fma(-0.125, (u2 * u2) * u2, 0.25 * ((t1 * t1) * t1));
    }
    /// f and df for q <= 2 (the caller's support test guarantees it up to rounding)
    __device__ static __forceinline__ void f_df(f64 q, f64 &f, f64 &df) {
        f64 t1 = 2. - q, t2 = 1. - q;
        f64 u2 = t2 + fabs(t2);
        f64 s1 = t1 * t1, s2 = u2 * u2;
        f  = fma(-0.125, s2 * u2, 0.25 * (s1 * t1));
        df = 0.75 * (s2 - s1);
    }
};
template<>
struct FastK<KM6> {
    __device__ static __forceinline__ f64 df(f64 q) {
        f64 t1 = 3. - q, t2 = 2. - q, t3 = 1. - q;
        f64 u1 = t1 + fabs(t1), u2 = t2 + fabs(t2), u3 = t3 + fabs(t3);
        f64 s1 = u1 * u1, s2 = u2 * u2, s3 = u3 * u3;
        // -5 (x1^4 - 6 x2^4 + 15 x3^4) with x = u / 2
        return -0.3125 * fma(15., s3 * s3, fma(-6., s2 * s2, s1 * s1));
    }
    /// f for q <= 3
    __device__ static __forceinline__ f64 f_in(f64 q) {
        f64 f, df;
        f_df(q, f, df);
        return f;
    }
    __device__ static __forceinline__ void f_df(f64 q, f64 &f, f64 &df) {
        f64 t1 = 3. - q, t2 = 2. - q, t3 = 1. - q;
        f64 u1 = t1 + fabs(t1), u2 = t2 + fabs(t2), u3 = t3 + fabs(t3);
        f64 s1 = u1 * u1, s2 = u2 * u2, s3 = u3 * u3;
        f64 q1 = s1 * s1, q2 = s2 * s2, q3 = s3 * s3;
        f  = 0.03125 * fma(15., q3 * u3, fma(-6., q2 * u2, q1 * u1));
        df = -0.3125 * fma(15., q3, fma(-6., q2, q1));
    }
};

// ---- h Newton iteration (all sweeps) + Ω ------------------------------------------------------------
/// Σ_b f(q_ab) and Σ_b (3 f + q f') over the list of one particle, by the G lanes of its group
template<class K, int G>
__device__ __forceinline__ void density_sums(
    const RankCsr &c, const Pack4 *__restrict__ SA, u32 s0, u32 s1, int sub, const Pack4 &a, f64 h_a, f64 &sf,
    f64 &sg) {
    const f64 hinv = 1. / h_a;
    const f64 lim  = h_a * h_a * (K::Rkern * K::Rkern);
    f64 f_acc = 0, g_acc = 0;
    // two list entries per trip (independent loads and arithmetic chains), indices one trip ahead
    auto pair = [&](const Pack4 &b) {
        f64 dx = a.a - b.a, dy = a.b - b.b, dz = a.c - b.c;
        f64 r2 = dx * dx + dy * dy + dz * dz;
        if (r2 <= lim) {
            f64 x = r2 + 1e-280; // the particle itself: r2 = 0
            f64 q = (x * fast_rsqrt(x)) * hinv;
            f64 f, df;
            FastK<K>::f_df(q, f, df);
            f_acc += f;
            g_acc += fma(q, df, 3 * f);
        }
    };
    u32 j   = s0 + sub;
    u32 rb0 = j < s1 ? c.list[j] : 0u;
    u32 rb1 = j + G < s1 ? c.list[j + G] : 0u;
    while (j + G < s1) {
        const u32 jn  = j + 2 * G;
        const u32 rn0 = jn < s1 ? c.list[jn] : 0u;
        const u32 rn1 = jn + G < s1 ? c.list[jn + G] : 0u;
        const Pack4 b0 = ld4(SA + rb0), b1 = ld4(SA + rb1);
        j   = jn;
        rb0 = rn0;
        rb1 = rn1;
        pair(b0);
        pair(b1);
    }
    if (j < s1)
        pair(ld4(SA + rb0));
    sf = group_sum<G>(f_acc);
    sg = group_sum<G>(g_acc);
}

template<class K, int G>
__global__ void __launch_bounds__(BLK) h_solve_fast_kernel(
    RankCsr c, const Pack4 *__restrict__ SA, const f64 *__restrict__ h_old, f64 *__restrict__ hpart,
    f64 *__restrict__ eps, f64 *__restrict__ omega, f64 part_mass, f64 h_max_tot_max_evol, f64 h_max_evol_p,
    u32 max_sweeps, bool do_iter, bool do_omega, u64 *red) {
    const u32 t   = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 k   = t / G;
    const int sub = int(t % G);
    // groups past the end clamp to the last particle (their lanes must stay in the shuffles) and write nothing
    const bool valid = k < c.count;
    u32 kk, r, id;
    csr_item(c, valid ? k : c.count - 1, kk, r, id);
    const Pack4 a    = ld4(SA + r);
    f64 h_a          = hpart[id];
    const u32 s0 = c.off[kk], s1 = s0 + c.cnt[kk];
    constexpr f64 hf3 = K::hfactd * K::hfactd * K::hfactd;
    f64 e_out  = 0;
    u32 sweeps = 0;
    f64 growth = 0; // largest h_iterate / h_old of this particle (red[6]: how much list tolerance the step needed)
    if (do_iter) {
        f64 e            = eps[id];
        const f64 ha_0   = h_old[id];
        f64 h_top        = h_a;
        const f64 h_max_evol_m = 1 / h_max_evol_p;
        while (sweeps < max_sweeps && e > 1e-6) {
            f64 sf, sg;
            density_sums<K, G>(c, SA, s0, s1, sub, a, h_a, sf, sg);
            f64 hinv    = 1. / h_a;
            f64 hinv3   = hinv * hinv * hinv;
            f64 rho_sum = part_mass * K::norm_3d * hinv3 * sf;
            f64 sumdWdh = -part_mass * K::norm_3d * hinv3 * hinv * sg;
            f64 rho_ha  = part_mass * hf3 * hinv3;
            f64 f_iter  = rho_sum - rho_ha;
            f64 df_iter = sumdWdh + 3 * rho_ha * hinv;
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
            h_top = fmax(h_top, h_a);
            sweeps++;
        }
        e_out  = e;
        growth = h_top / ha_0;
        if (valid && sub == 0) {
            eps[id]   = e;
            hpart[id] = h_a;
        }
    }
    if (do_omega) {
        f64 sf, sg;
        density_sums<K, G>(c, SA, s0, s1, sub, a, h_a, sf, sg);
        // Ω = 1 + h/(3 ρ_h) Σ m ∂W/∂h = 1 - (norm / (3 hfact³)) Σ (3 f + q f')
        if (valid && sub == 0)
            omega[id] = 1 - (K::norm_3d / (3 * hf3)) * sg;
    }
    if (do_iter) {
        __shared__ f64 smax[BLK / 32], smin[BLK / 32], sgr[BLK / 32];
        __shared__ u32 ssw[BLK / 32];
        const bool mine = valid && sub == 0;
        f64 vgr  = warp_max(mine ? growth : 0.);
        f64 vmax = warp_max(mine ? e_out : -f64(INFINITY));
        f64 vmin = warp_min(mine ? e_out : f64(INFINITY));
        u32 sw   = mine ? sweeps : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            sw = max(sw, __shfl_xor_sync(0xffffffffu, sw, o));
        if ((threadIdx.x & 31) == 0) {
            smax[threadIdx.x >> 5] = vmax;
            smin[threadIdx.x >> 5] = vmin;
            sgr[threadIdx.x >> 5]  = vgr;
            ssw[threadIdx.x >> 5]  = sw;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            f64 x = smax[0], y = smin[0], g = sgr[0];
            u32 z = ssw[0];
#pragma unroll
            for (int q = 1; q < BLK / 32; q++) {
                x = fmax(x, smax[q]);
                y = fmin(y, smin[q]);
                g = fmax(g, sgr[q]);
                z = max(z, ssw[q]);
            }
            atomicMax((unsigned long long *) &red[6], (unsigned long long) f64_to_ordered(g));
            atomicMax((unsigned long long *) &red[0], (unsigned long long) f64_to_ordered(x));
            atomicMin((unsigned long long *) &red[1], (unsigned long long) f64_to_ordered(y));
            atomicMax((unsigned long long *) &red[2], (unsigned long long) z);
        }
    }
}

// ---- ∇·v, ∇×v, d(∇·v)/dt ------------------------------------------------------------------------------
/// OMEGA: the pass also accumulates Σ (3 f + q f') over the support of h_a (the Ω sum of h_solve, self pair
/// included) and writes Ω_a: one pass over the lists less per step
template<class K, int G, bool SPHDIV, bool CURL, bool MAT, bool COMBINED, bool OMEGA>
__global__ void __launch_bounds__(BLK) av_operators_fast_kernel(
    RankCsr c, const Pack4 *__restrict__ SA, const Pack4 *__restrict__ SB, const Pack4 *__restrict__ SC,
    const Pack4 *__restrict__ SD, f64 pmass, f64 *__restrict__ divv, f64 *__restrict__ curlv,
    f64 *__restrict__ dtdivv, f64 *__restrict__ omega_out) {
    const u32 t      = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 k      = t / G;
    const int sub    = int(t % G);
    const bool valid = k < c.count;
    u32 kk, r, id;
    csr_item(c, valid ? k : c.count - 1, kk, r, id);
    constexpr f64 Rker2 = K::Rkern * K::Rkern;
    const Pack4 pa = ld4(SA + r), va = ld4(SB + r);
    const Pack4 aa = MAT ? ld4(SD + r) : Pack4{0, 0, 0, 0};
    const f64 h_a   = pa.d;
    const f64 hinv  = 1. / h_a;
    const f64 dWn_a = K::norm_3d * (hinv * hinv) * (hinv * hinv);
    const f64 lim_a = h_a * h_a * Rker2;
    // MAT: Rij = -Σ r ⊗ ∇W is symmetric (∇W ∥ r): six sums; Rv = -Σ ∇W ⊗ v holds every product v_m ∂_i W, so
    // the SPH divergence and curl (Σ v·∇W, Σ v × ∇W) are read off it instead of being summed a second time
    constexpr bool OWN_DIV = SPHDIV && !MAT;
    f64 snv = 0, cx = 0, cy = 0, cz = 0, sg = 0;
    f64 S[6], Rv[9], Ra[9]; // S: xx xy xz yy yz zz
    if (MAT) {
#pragma unroll
        for (int i = 0; i < 9; i++) {
            Rv[i] = 0;
            Ra[i] = 0;
        }
#pragma unroll
        for (int i = 0; i < 6; i++)
            S[i] = 0;
    }
    const u32 s0 = c.off[kk], s1 = s0 + c.cnt[kk];
    u32 j  = s0 + sub;
    u32 rb = j < s1 ? c.list[j] : 0u;
    while (j < s1) { // next index one trip ahead, the records of this trip requested together
        const u32 jn  = j + G;
        const u32 rbn = jn < s1 ? c.list[jn] : rb;
        const Pack4 pb = ld4(SA + rb), vb = ld4(SB + rb);
        const Pack4 ab = MAT ? ld4(SD + rb) : Pack4{0, 0, 0, 0};
        j  = jn;
        rb = rbn;
        f64 dx = pa.a - pb.a, dy = pa.b - pb.b, dz = pa.c - pb.c;
        f64 r2  = dx * dx + dy * dy + dz * dz;
        f64 h_b = pb.d;
        if (r2 > lim_a && r2 > h_b * h_b * Rker2)
            continue;
        f64 x    = OMEGA ? r2 + 1e-280 : r2; // the particle itself (r2 = 0) counts in the Ω sum
        f64 rinv = fast_rsqrt(x);
        f64 q    = (x * rinv) * hinv;
        f64 dfq  = FastK<K>::df(q);
        if (OMEGA && r2 <= lim_a)
            sg += fma(q, dfq, 3. * FastK<K>::f_in(q));
        if (r2 < 1e-18) // r < 1e-9: zero unit vector in the reference
            continue;
        f64 gs   = dWn_a * dfq * rinv; // ∇W_ab(h_a) = gs · r_ab  (mass factored out)
        f64 gx = gs * dx, gy = gs * dy, gz = gs * dz;
        f64 vx = va.a - vb.a, vy = va.b - vb.b, vz = va.c - vb.c;
        if (OWN_DIV) {
            snv += vx * gx + vy * gy + vz * gz;
            if (CURL) {
                cx += vy * gz - vz * gy;
                cy += vz * gx - vx * gz;
                cz += vx * gy - vy * gx;
            }
        }
        if (MAT) {
            f64 v[3] = {vx, vy, vz};
            f64 a[3] = {aa.a - ab.a, aa.b - ab.b, aa.c - ab.c};
            f64 g[3] = {gx, gy, gz};
            S[0] -= dx * gx, S[1] -= dx * gy, S[2] -= dx * gz;
            S[3] -= dy * gy, S[4] -= dy * gz, S[5] -= dz * gz;
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int m = 0; m < 3; m++) {
                    Rv[3 * i + m] -= v[m] * g[i];
                    Ra[3 * i + m] -= a[m] * g[i];
                }
        }
    }
    f64 Rij[9];
    if (OWN_DIV) {
        snv = group_sum<G>(snv);
        if (CURL) {
            cx = group_sum<G>(cx);
            cy = group_sum<G>(cy);
            cz = group_sum<G>(cz);
        }
    }
    if (MAT) {
#pragma unroll
        for (int i = 0; i < 9; i++) {
            Rv[i] = group_sum<G>(Rv[i]);
            Ra[i] = group_sum<G>(Ra[i]);
        }
#pragma unroll
        for (int i = 0; i < 6; i++)
            S[i] = group_sum<G>(S[i]);
        Rij[0] = S[0], Rij[1] = S[1], Rij[2] = S[2], Rij[3] = S[1], Rij[4] = S[3], Rij[5] = S[4], Rij[6] = S[2],
        Rij[7] = S[4], Rij[8] = S[5];
        if (SPHDIV) { // Rv[3 i + m] = -Σ v_m ∂_i W
            snv = -(Rv[0] + Rv[4] + Rv[8]);
            cx  = Rv[5] - Rv[7];
            cy  = Rv[6] - Rv[2];
            cz  = Rv[1] - Rv[3];
        }
    }
    f64 omega_new = 0;
    if (OMEGA) { // Ω = 1 + h/(3 ρ_h) Σ m ∂W/∂h = 1 - (norm / (3 hfact³)) Σ (3 f + q f')   (h_solve_fast_kernel)
        constexpr f64 hf3 = K::hfactd * K::hfactd * K::hfactd;
        omega_new         = 1 - (K::norm_3d / (3 * hf3)) * group_sum<G>(sg);
    }
    if (!valid || sub != 0)
        return;
    if (OMEGA)
        omega_out[id] = omega_new;
    if (SPHDIV) {
        f64 omega_a = OMEGA ? omega_new : SC[r].b;
        f64 hfh     = K::hfactd * hinv;
        f64 rho_a   = pmass * hfh * hfh * hfh;
        f64 fac     = -pmass / (omega_a * rho_a);
        divv[id]    = fac * snv;
        if (CURL) {
            curlv[3 * u64(id)]     = fac * cx;
            curlv[3 * u64(id) + 1] = fac * cy;
            curlv[3 * u64(id) + 2] = fac * cz;
        }
    }
    if (MAT) { // dv = Rij^-1 Rv, da = Rij^-1 Ra  (matrix_legacy.hpp:25,53); the mass cancels
        f64 a00 = Rij[0], a01 = Rij[1], a02 = Rij[2], a10 = Rij[3], a11 = Rij[4], a12 = Rij[5], a20 = Rij[6],
            a21 = Rij[7], a22 = Rij[8];
        f64 det = (-a02 * a11 * a20 + a01 * a12 * a20 + a02 * a10 * a21 - a00 * a12 * a21 - a01 * a10 * a22
                   + a00 * a11 * a22);
        f64 idet = 1. / det;
        f64 inv[9];
        inv[0] = (-a12 * a21 + a11 * a22) * idet;
        inv[1] = (a02 * a21 - a01 * a22) * idet;
        inv[2] = (-a02 * a11 + a01 * a12) * idet;
        inv[3] = (a12 * a20 - a10 * a22) * idet;
        inv[4] = (-a02 * a20 + a00 * a22) * idet;
        inv[5] = (a02 * a10 - a00 * a12) * idet;
        inv[6] = (-a11 * a20 + a10 * a21) * idet;
        inv[7] = (a01 * a20 - a00 * a21) * idet;
        inv[8] = (-a01 * a10 + a00 * a11) * idet;
        f64 dv[9], da_tr = 0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
#pragma unroll
            for (int m = 0; m < 3; m++)
                dv[3 * i + m] = inv[3 * i] * Rv[m] + inv[3 * i + 1] * Rv[3 + m] + inv[3 * i + 2] * Rv[6 + m];
            da_tr += inv[3 * i] * Ra[i] + inv[3 * i + 1] * Ra[3 + i] + inv[3 * i + 2] * Ra[6 + i];
        }
        f64 tens = 0;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int m = 0; m < 3; m++)
                tens += dv[3 * m + i] * dv[3 * i + m];
        if (COMBINED) {
            divv[id]               = dv[0] + dv[4] + dv[8];
            curlv[3 * u64(id)]     = dv[5] - dv[7];
            curlv[3 * u64(id) + 1] = dv[6] - dv[2];
            curlv[3 * u64(id) + 2] = dv[1] - dv[3];
        }
        dtdivv[id] = da_tr - tens;
    }
}

// ---- per-particle derived factors ----------------------------------------------------------------------
/// SE = (x, y, z, 1/h), SF = (1/(rho² Ω), α c_s, P, rho): with SB = (v, u) the force loop reads THREE
/// 32-byte records per neighbour (the loop is bound by L1 tag lookups, one per record and line)
template<class K>
__global__ void __launch_bounds__(256) derive_fast_kernel(
    u32 M, bool vary, f64 alpha_const, const Pack4 *__restrict__ SA, const Pack4 *__restrict__ SC, f64 pmass,
    Pack4 *__restrict__ SE, Pack4 *__restrict__ SF, double2 *__restrict__ SG) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M)
        return;
    Pack4 pa = SA[i];
    Pack4 cc = SC[i]; // (P, omega, cs, alpha)
    f64 hinv = 1. / pa.d;
    f64 hfh  = K::hfactd * hinv;
    f64 rho  = pmass * hfh * hfh * hfh;
    f64 sub  = rho * rho * cc.b;
    f64 iro2 = (sub != 0. && sub == sub) ? 1. / sub : 0.; // inv_sat_zero
    f64 alpha = vary ? cc.d : alpha_const;
    SE[i] = Pack4{pa.a, pa.b, pa.c, hinv};
    SF[i] = Pack4{iro2, alpha * cc.c, cc.a, rho};
    if (SG)
        SG[i] = make_double2(iro2, alpha * cc.c);
}

// ---- forces + v_sig + CFL --------------------------------------------------------------------------------
template<class K, int AV, int G, bool SF16 = false>
__global__ void __launch_bounds__(BLK) force_cfl_fast_kernel(
    RankCsr c, const Pack4 *__restrict__ SA, const Pack4 *__restrict__ SB, const Pack4 *__restrict__ SE,
    const Pack4 *__restrict__ SF, const Pack4 *__restrict__ SC, SphParams p, const f64 *__restrict__ axyz_ext,
    f64 *__restrict__ axyz, f64 *__restrict__ duint, f64 C_cour, f64 C_force, f64 *__restrict__ vsig_out,
    f64 *__restrict__ cfl_out, u64 *red_min) {
    const u32 t      = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 k      = t / G;
    const int sub    = int(t % G);
    const bool valid = k < c.count;
    u32 kk, r, id;
    csr_item(c, valid ? k : c.count - 1, kk, r, id);
    constexpr f64 Rker2 = K::Rkern * K::Rkern;
    constexpr bool DISC = (AV == AVK_DISC);
    const Pack4 pa = ld4(SE + r), va = ld4(SB + r), fa = ld4(SF + r);
    const f64 u_a = va.d;
    const f64 hinv_a = pa.d, h_a = SA[r].d;
    const f64 iro2_a = fa.a, acs_a = fa.b, P_a = fa.c, rho_a = fa.d;
    const f64 dWn_a  = K::norm_3d * (hinv_a * hinv_a) * (hinv_a * hinv_a);
    const f64 iro_a  = iro2_a * rho_a;
    const f64 Pfac_a = P_a * iro2_a;
    const f64 cs_a   = SC[r].c;
    const f64 lim_a  = h_a * h_a * Rker2;
    f64 fx = 0, fy = 0, fz = 0, dU1 = 0, dU2 = 0, vsig_max = 0;
    const u32 s0 = c.off[kk], s1 = s0 + c.cnt[kk];
    // software pipeline: the index of the next neighbour is fetched one trip ahead and the three records of
    // the current one are requested together (one 256-bit load each), so a trip waits for ONE memory latency
    u32 j  = s0 + sub;
    u32 rb = j < s1 ? c.list[j] : 0u;
    while (j < s1) {
        const u32 jn  = j + G;
        const u32 rbn = jn < s1 ? c.list[jn] : rb;
        const Pack4 pb = ld4(SE + rb), vb = ld4(SB + rb);
        Pack4 fb;
        if (SF16) { // 16 bytes instead of 32: rho_b from h_b, P_b = (gamma - 1) rho_b u_b
            double2 g;
            asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(g.x), "=d"(g.y) : "l"(p.SG + rb));
            const f64 hfh = K::hfactd * pb.d;
            const f64 rho = p.pmass * hfh * hfh * hfh;
            fb            = Pack4{g.x, g.y, p.adiabatic_gm1 * rho * vb.d, rho};
        } else {
            fb = ld4(SF + rb);
        }
        j  = jn;
        rb = rbn;
        f64 dx = pa.a - pb.a, dy = pa.b - pb.b, dz = pa.c - pb.c;
        f64 r2     = dx * dx + dy * dy + dz * dz;
        f64 hinv_b = pb.d;
        if (r2 > lim_a && r2 * (hinv_b * hinv_b) > Rker2)
            continue;
        if (r2 < 1e-18) { // r < 1e-9 (the particle itself): zero unit vector, only v_sig sees the pair
            vsig_max = fmax(vsig_max, cs_a);
            continue;
        }
        f64 rinv = fast_rsqrt(r2);
        f64 rab  = r2 * rinv;
        f64 vx = va.a - vb.a, vy = va.b - vb.b, vz = va.c - vb.c;
        f64 vr  = (vx * dx + vy * dy + vz * dz) * rinv;
        f64 avr = fabs(vr);
        f64 hb2 = hinv_b * hinv_b;
        f64 Fa  = dWn_a * FastK<K>::df(rab * hinv_a);
        f64 Fb  = (K::norm_3d * FastK<K>::df(rab * hinv_b)) * (hb2 * hb2);
        f64 vsig_a = acs_a + p.beta_AV * avr;
        f64 vsig_b = fb.b + p.beta_AV * avr;
        f64 rho_b  = fb.d;
        f64 qa_ab, qb_ab;
        if (DISC) { // q_av_disc (q_ab.hpp:42-60)
            f64 vd_a = (vr < 0.) ? vsig_a : acs_a;
            f64 vd_b = (vr < 0.) ? vsig_b : fb.b;
            qa_ab    = (-0.5 * rho_a * rinv * h_a) * vd_a * vr;
            qb_ab    = (-0.5 * rho_b * rinv * __drcp_rn(hinv_b)) * vd_b * vr;
        } else { // q_av (q_ab.hpp:37-40)
            // max(x, 0) = (x + |x|) / 2 with x = -rho vsig vr / 2: one add instead of a compare + two selects
            f64 xa = (rho_a * vsig_a) * vr, xb = (rho_b * vsig_b) * vr;
            qa_ab  = 0.25 * (fabs(xa) - xa);
            qb_ab  = 0.25 * (fabs(xb) - xb);
        }
        f64 ka = Pfac_a + qa_ab * iro2_a; // (P_a + q_a) / (rho_a² Ω_a)
        f64 kb = (fb.c + qb_ab) * fb.a;
        f64 cf = (ka * Fa + kb * Fb) * rinv;
        fx += cf * dx;
        fy += cf * dy;
        fz += cf * dz;
        dU1 += ka * vr * Fa;
        // sqrt(A / B) = A rsqrt(A B), A = 2 |P_a - P_b|, B = rho_a + rho_b (A = 0: 0 * rsqrt(1e-280) = 0)
        f64 Apr    = 2. * fabs(P_a - fb.c);
        f64 vsig_u = Apr * fast_rsqrt(fma(Apr, rho_a + rho_b, 1e-280));
        dU2 += vsig_u * (u_a - vb.d) * (Fa * iro_a + Fb * (fb.a * rho_b));
        vsig_max = fmax(vsig_max, cs_a + 2.0 * avr);
    }
    fx       = group_sum<G>(fx);
    fy       = group_sum<G>(fy);
    fz       = group_sum<G>(fz);
    dU1      = group_sum<G>(dU1);
    dU2      = group_sum<G>(dU2);
    vsig_max = group_max<G>(vsig_max);
    f64 dt_out = f64(INFINITY);
    if (valid && sub == 0) {
        f64 ax = -p.pmass * fx + axyz_ext[3 * u64(id)];
        f64 ay = -p.pmass * fy + axyz_ext[3 * u64(id) + 1];
        f64 az = -p.pmass * fz + axyz_ext[3 * u64(id) + 2];
        axyz[3 * u64(id)]     = ax;
        axyz[3 * u64(id) + 1] = ay;
        axyz[3 * u64(id) + 2] = az;
        duint[id]             = p.pmass * (dU1 + 0.5 * p.alpha_u * dU2);
        vsig_out[id]          = vsig_max;
        f64 dt_c = C_cour * h_a / vsig_max;
        f64 dt_f = C_force * sqrt(h_a / sqrt(ax * ax + ay * ay + az * az));
        dt_out   = fmin(dt_c, dt_f);
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

template<int G>
unsigned grid_groups(u32 N) {
    return grid_for(u64(N) * G, BLK);
}

} // namespace

// lanes per particle.  Default: measured optimum per kernel shape (profiles/); SHAMB200_LANES overrides
// (1, 2, 4, 8, 16, 32) for tuning runs.
static int lanes_override() {
    static int v = [] {
        const char *e = getenv("SHAMB200_LANES");
        return e ? atoi(e) : 0;
    }();
    return v;
}
#define SB_G_CASE(GV, CALL)                                                                      \
    case GV: {                                                                                   \
        constexpr int G = GV;                                                                    \
        CALL;                                                                                    \
    } break;
#define SB_KDG(kernel, CALL)                                                                     \
    do {                                                                                         \
        int g_ = lanes_override();                                                               \
        if ((kernel) == KERN_M4) {                                                               \
            using KT = KM4;                                                                      \
            switch (g_ ? g_ : 2) {                                                               \
                SB_G_CASE(1, CALL) SB_G_CASE(2, CALL) SB_G_CASE(4, CALL) SB_G_CASE(8, CALL)      \
                SB_G_CASE(16, CALL) SB_G_CASE(32, CALL)                                          \
            default: throw std::invalid_argument("SHAMB200_LANES must be 1, 2, 4, 8, 16 or 32"); \
            }                                                                                    \
        } else {                                                                                 \
            using KT = KM6;                                                                      \
            switch (g_ ? g_ : 4) {                                                               \
                SB_G_CASE(1, CALL) SB_G_CASE(2, CALL) SB_G_CASE(4, CALL) SB_G_CASE(8, CALL)      \
                SB_G_CASE(16, CALL) SB_G_CASE(32, CALL)                                          \
            default: throw std::invalid_argument("SHAMB200_LANES must be 1, 2, 4, 8, 16 or 32"); \
            }                                                                                    \
        }                                                                                        \
        SB_COUNT_LAUNCH();                                                                       \
        SB_LAUNCH_CHECK();                                                                       \
    } while (0)

void h_solve_fast(
    cudaStream_t s, int kernel, RankCsr c, const Pack4 *SA, const f64 *h_old, f64 *hpart, f64 *eps, f64 *omega,
    f64 pmass, f64 h_evol_max, f64 h_evol_iter_max, u32 max_sweeps, bool do_iter, bool do_omega, u64 *red) {
    if (!c.N || !c.count)
        return;
    SB_KDG(kernel, (h_solve_fast_kernel<KT, G><<<grid_groups<G>(c.count), BLK, 0, s>>>(
                       c, SA, h_old, hpart, eps, omega, pmass, h_evol_max, h_evol_iter_max, max_sweeps, do_iter,
                       do_omega, red)));
}

void av_operators_fast(
    cudaStream_t s, int kernel, RankCsr c, const Pack4 *SA, const Pack4 *SB, const Pack4 *SC, const Pack4 *SD,
    f64 pmass, bool want_curl, bool want_dtdivv, bool combined, f64 *divv, f64 *curlv, f64 *dtdivv, f64 *omega_out) {
    if (!c.N || !c.count)
        return;
#define AVOP2(S_, C_, M_, CB_, OM_)                                                              \
    SB_KDG(kernel, (av_operators_fast_kernel<KT, G, S_, C_, M_, CB_, OM_><<<grid_groups<G>(c.count), BLK, 0, s>>>( \
                       c, SA, SB, SC, SD, pmass, divv, curlv, dtdivv, omega_out)))
#define AVOP(S_, C_, M_, CB_)                                                                    \
    do {                                                                                         \
        if (omega_out)                                                                           \
            AVOP2(S_, C_, M_, CB_, true);                                                        \
        else                                                                                     \
            AVOP2(S_, C_, M_, CB_, false);                                                       \
    } while (0)
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
#undef AVOP2
}

void derive_fast(
    cudaStream_t s, int kernel, int av, u32 M, const Pack4 *SA, const Pack4 *SB, const Pack4 *SC, f64 pmass,
    f64 alpha_AV, Pack4 *SE, Pack4 *SF, double2 *SG) {
    (void) SB;
    if (!M)
        return;
    bool vary = (av == AVK_MM97 || av == AVK_CD10);
    if (kernel == KERN_M4)
        derive_fast_kernel<KM4><<<grid_for(M, 256), 256, 0, s>>>(M, vary, alpha_AV, SA, SC, pmass, SE, SF, SG);
    else
        derive_fast_kernel<KM6><<<grid_for(M, 256), 256, 0, s>>>(M, vary, alpha_AV, SA, SC, pmass, SE, SF, SG);
    SB_COUNT_LAUNCH();
    SB_LAUNCH_CHECK();
}

void force_cfl_fast(
    cudaStream_t s, int kernel, int av, RankCsr c, const Pack4 *SA, const Pack4 *SB, const Pack4 *SE, const Pack4 *SF,
    const Pack4 *SC, SphParams p, const f64 *axyz_ext, f64 *axyz, f64 *duint, f64 C_cour, f64 C_force, f64 *vsig,
    f64 *cfl_dt, u64 *red_min) {
    if (!c.N || !c.count)
        return;
#define FRC(AV_)                                                                                 \
    do {                                                                                         \
        if (p.SG && p.adiabatic_gm1 != 0)                                                        \
            SB_KDG(kernel, (force_cfl_fast_kernel<KT, AV_, G, true><<<grid_groups<G>(c.count), BLK, 0, s>>>(  \
                               c, SA, SB, SE, SF, SC, p, axyz_ext, axyz, duint, C_cour, C_force, vsig, cfl_dt, red_min))); \
        else                                                                                     \
            SB_KDG(kernel, (force_cfl_fast_kernel<KT, AV_, G><<<grid_groups<G>(c.count), BLK, 0, s>>>(   \
                               c, SA, SB, SE, SF, SC, p, axyz_ext, axyz, duint, C_cour, C_force, vsig, cfl_dt, red_min))); \
    } while (0)
    switch (av) {
    case AVK_CONSTANT:
    case AVK_MM97:
    case AVK_CD10: FRC(AVK_CD10); break; // α·c_s comes from the derived record in every non-disc mode
    case AVK_DISC: FRC(AVK_DISC); break;
    default: throw std::invalid_argument("unsupported artificial viscosity configuration");
    }
#undef FRC
}

} // namespace sb
