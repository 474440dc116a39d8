// sph2.cu — fp_mode dispatch of the model-path SPH loops (sph2.cuh)
#include "sph2.cuh"

namespace sb {

void h_solve(
    cudaStream_t s, int fp_mode, int kernel, RankCsr c, const Pack4 *SA, const f64 *h_old, f64 *hpart, f64 *eps,
    f64 *omega, f64 pmass, f64 h_evol_max, f64 h_evol_iter_max, u32 max_sweeps, bool do_iter, bool do_omega,
    u64 *red) {
    if (fp_mode == FP_FAST)
        h_solve_fast(s, kernel, c, SA, h_old, hpart, eps, omega, pmass, h_evol_max, h_evol_iter_max, max_sweeps, do_iter, do_omega, red);
    else
        h_solve_strict(s, kernel, c, SA, h_old, hpart, eps, omega, pmass, h_evol_max, h_evol_iter_max, max_sweeps, do_iter, do_omega, red);
}

void av_operators(
    cudaStream_t s, int fp_mode, int kernel, RankCsr c, const Pack4 *SA, const Pack4 *SB, const Pack4 *SC,
    const Pack4 *SD, f64 pmass, bool want_curl, bool want_dtdivv, bool combined, f64 *divv, f64 *curlv, f64 *dtdivv,
    f64 *omega_out) {
    if (fp_mode == FP_FAST)
        av_operators_fast(s, kernel, c, SA, SB, SC, SD, pmass, want_curl, want_dtdivv, combined, divv, curlv, dtdivv,
                          omega_out);
    else
        av_operators_strict(s, kernel, c, SA, SB, SC, SD, pmass, want_curl, want_dtdivv, combined, divv, curlv, dtdivv);
}

void force_cfl(
    cudaStream_t s, int fp_mode, int kernel, int av, RankCsr c, const Pack4 *SA, const Pack4 *SB, const Pack4 *SC,
    const Pack4 *SE, const Pack4 *SF, SphParams p, const f64 *axyz_ext, f64 *axyz, f64 *duint, f64 C_cour,
    f64 C_force, f64 *vsig, f64 *cfl_dt, u64 *red_min) {
    if (fp_mode == FP_FAST)
        force_cfl_fast(s, kernel, av, c, SA, SB, SE, SF, SC, p, axyz_ext, axyz, duint, C_cour, C_force, vsig, cfl_dt, red_min);
    else
        force_cfl_strict(s, kernel, av, c, SA, SB, SC, p, axyz_ext, axyz, duint, C_cour, C_force, vsig, cfl_dt, red_min);
}

} // namespace sb
