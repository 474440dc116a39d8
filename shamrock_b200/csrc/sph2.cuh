// sph2.cuh — the SPH neighbour loops of the model path, on Morton-sorted records and the rank CSR of
// neigh2.cuh.  Three passes over the neighbour lists per step:
//   h_solve       all Newton sweeps of the smoothing length (each particle iterates in-kernel: a sweep
//                 only reads the particle's own h, so this equals the reference's synchronous sweeps)
//                 + Ω with the converged h                                  (K20 + K21 + K22)
//   av_operators  ∇·v, ∇×v and d(∇·v)/dt in one loop                         (K26 + K27)
//   force_cfl     dv/dt, du/dt, v_sig and the CFL dt with its min reduction   (K29 + K30 + K33 + K34)
// fp_mode 0 (strict, sph2_strict.cu, -fmad=false): thread per particle, the reference's expressions in
// the reference's order → bit-identical to the oracle.  fp_mode 1 (fast, sph2_fast.cu, FMA allowed):
// several lanes per particle over the list, per-particle reciprocals precomputed, rsqrt-based pair
// math → within 1e-10 relative.
#pragma once
#include "neigh2.cuh"
#include "sph.cuh"

namespace sb {

enum { FP_STRICT = 0, FP_FAST = 1 };

/// red: [0] max eps (ordered), [1] min eps (ordered), [2] max sweep count (u64); fast fp mode also [6] = max over
/// the particles and sweeps of h_iterate / h_old (ordered): the list tolerance this iteration needed (solver.cu)
void h_solve(
    cudaStream_t s, int fp_mode, int kernel, RankCsr c, const Pack4 *SA, const f64 *h_old, f64 *hpart, f64 *eps,
    f64 *omega, f64 pmass, f64 h_evol_max, f64 h_evol_iter_max, u32 max_sweeps, bool do_iter, bool do_omega,
    u64 *red);

/// omega_out (fast fp mode only): the pass also sums Ω of every real particle (ComputeOmega.cpp:36-73) with the
/// converged h and uses it for its own 1 / (ρ Ω) factor — h_solve is then run without its Ω pass
void av_operators(
    cudaStream_t s, int fp_mode, int kernel, RankCsr c, const Pack4 *SA, const Pack4 *SB, const Pack4 *SC,
    const Pack4 *SD, f64 pmass, bool want_curl, bool want_dtdivv, bool combined, f64 *divv, f64 *curlv, f64 *dtdivv,
    f64 *omega_out = nullptr);

void force_cfl(
    cudaStream_t s, int fp_mode, int kernel, int av, RankCsr c, const Pack4 *SA, const Pack4 *SB, const Pack4 *SC,
    const Pack4 *SE, const Pack4 *SF, SphParams p, const f64 *axyz_ext, f64 *axyz, f64 *duint, f64 C_cour, f64 C_force, f64 *vsig, f64 *cfl_dt,
    u64 *red_min);

// strict / fast implementations (one translation unit each)
void h_solve_strict(
    cudaStream_t s, int kernel, RankCsr c, const Pack4 *SA, const f64 *h_old, f64 *hpart, f64 *eps, f64 *omega,
    f64 pmass, f64 h_evol_max, f64 h_evol_iter_max, u32 max_sweeps, bool do_iter, bool do_omega, u64 *red);
void av_operators_strict(
    cudaStream_t s, int kernel, RankCsr c, const Pack4 *SA, const Pack4 *SB, const Pack4 *SC, const Pack4 *SD,
    f64 pmass, bool want_curl, bool want_dtdivv, bool combined, f64 *divv, f64 *curlv, f64 *dtdivv);
void force_cfl_strict(
    cudaStream_t s, int kernel, int av, RankCsr c, const Pack4 *SA, const Pack4 *SB, const Pack4 *SC, SphParams p,
    const f64 *axyz_ext, f64 *axyz, f64 *duint, f64 C_cour, f64 C_force, f64 *vsig, f64 *cfl_dt, u64 *red_min);
void h_solve_fast(
    cudaStream_t s, int kernel, RankCsr c, const Pack4 *SA, const f64 *h_old, f64 *hpart, f64 *eps, f64 *omega,
    f64 pmass, f64 h_evol_max, f64 h_evol_iter_max, u32 max_sweeps, bool do_iter, bool do_omega, u64 *red);
void av_operators_fast(
    cudaStream_t s, int kernel, RankCsr c, const Pack4 *SA, const Pack4 *SB, const Pack4 *SC, const Pack4 *SD,
    f64 pmass, bool want_curl, bool want_dtdivv, bool combined, f64 *divv, f64 *curlv, f64 *dtdivv,
    f64 *omega_out = nullptr);
/// SE / SF: per-particle derived factors written by derive_fast (merged range, sorted order)
void derive_fast(
    cudaStream_t s, int kernel, int av, u32 M, const Pack4 *SA, const Pack4 *SB, const Pack4 *SC, f64 pmass,
    f64 alpha_AV, Pack4 *SE, Pack4 *SF, double2 *SG = nullptr);
void force_cfl_fast(
    cudaStream_t s, int kernel, int av, RankCsr c, const Pack4 *SA, const Pack4 *SB, const Pack4 *SE, const Pack4 *SF,
    const Pack4 *SC, SphParams p,
    const f64 *axyz_ext, f64 *axyz, f64 *duint, f64 C_cour, f64 C_force, f64 *vsig, f64 *cfl_dt, u64 *red_min);

} // namespace sb
