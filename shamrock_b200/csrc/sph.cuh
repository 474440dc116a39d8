// sph.cuh — host-callable launchers of the SPH loops (defined in sph.cu).
#pragma once
#include "common.cuh"
#include "neigh.cuh"

namespace sb {

/// 32-byte records used on the gather side of the neighbour loops (one sector each)
struct alignas(32) Pack4 {
    f64 a, b, c, d;
};

#ifdef __CUDACC__
/// one 32-byte record = ONE 256-bit read-only load (LDG.E.256 on sm_100a).  The neighbour loops are bound
/// by L1 tag lookups (one per distinct 128-byte line per load instruction), so a record must not be
/// fetched as two 16-byte halves.
__device__ __forceinline__ Pack4 ld4(const Pack4 *p) {
    Pack4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
    return r;
}
#endif

struct CsrView {
    const u32 *cnt, *scanned, *list;
    u32 N;
};

enum { KERN_M4 = 0, KERN_M6 = 1 };
enum { AVK_CONSTANT = 1, AVK_MM97 = 2, AVK_CD10 = 3, AVK_DISC = 4 };
enum { EOSK_ADIABATIC = 0, EOSK_ISOTHERMAL = 1, EOSK_LP07 = 2 };

struct SphParams {
    f64 pmass;
    f64 alpha_u, alpha_AV, beta_AV;
    /// fast force loop, adiabatic EOS: gamma - 1 and the 16-byte records (1/(rho² Ω), α c_s) of derive_fast; the
    /// neighbour's rho and P are then recomputed from its h and u instead of being loaded (0 / null: not used)
    f64 adiabatic_gm1   = 0;
    const double2 *SG   = nullptr;
};

/// One Newton sweep (IterateSmoothingLengthDensity).  order: thread→object map (may be null),
/// n_order its length.  red: 2 u64 (ordered-encoded running max, min of eps), accumulated.
void h_iterate(
    cudaStream_t s, int kernel, CsrView c, const f64 *xyz, size_t stride, const u32 *order, u32 n_order,
    const f64 *h_old, f64 *h_new, f64 *eps, f64 pmass, f64 h_evol_max, f64 h_evol_iter_max, u64 *red);
void compute_omega(
    cudaStream_t s, int kernel, CsrView c, const f64 *xyz, size_t stride, const u32 *order, u32 n_order,
    const f64 *hpart, f64 *omega, f64 pmass);

/// EOS over the merged range (real + ghosts): C.a = P, C.c = cs
void compute_eos(
    cudaStream_t s, int kernel, int eos, const Pack4 *A, const Pack4 *B, Pack4 *C, u32 M, f64 pmass,
    f64 gamma, f64 cs0, f64 q, f64 r0);

void compute_divv_curlv(
    cudaStream_t s, int kernel, CsrView c, const Pack4 *A, const Pack4 *B, const Pack4 *C, const u32 *order,
    u32 n_order, f64 pmass, f64 *divv, f64 *curlv /*may be null*/);
void compute_dtdivv(
    cudaStream_t s, int kernel, CsrView c, const Pack4 *A, const Pack4 *B, const Pack4 *D, const u32 *order,
    u32 n_order, f64 pmass, bool also_div_curl, f64 *divv, f64 *curlv, f64 *dtdivv);
void update_av(
    cudaStream_t s, int av, u32 N, f64 dt, f64 sigma_decay, f64 alpha_min, f64 alpha_max, const f64 *divv,
    const f64 *curlv, const f64 *dtdivv, const f64 *cs, const f64 *h, const f64 *alpha, f64 *alpha_updated);
/// axyz = force + axyz_ext ; duint
void compute_forces(
    cudaStream_t s, int kernel, int av, CsrView c, const Pack4 *A, const Pack4 *B, const Pack4 *C,
    const u32 *order, u32 n_order, SphParams p, const f64 *axyz_ext, f64 *axyz, f64 *duint);
/// vsig then cfl dt (Courant + force); red_min: ordered-encoded running min of the cfl dt
void compute_vsig_cfl(
    cudaStream_t s, int kernel, CsrView c, const Pack4 *A, const Pack4 *B, const Pack4 *C, const u32 *order,
    u32 n_order, const f64 *axyz, f64 C_cour, f64 C_force, f64 *vsig, f64 *cfl_dt, u64 *red_min);

} // namespace sb
