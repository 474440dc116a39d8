#!/usr/bin/env python
"""3D Sod shock tube on the B200 backend, driven through the reference's Python surface
(`shamrock_b200.pyshamrock`): the same sequence of calls as the reference's CI case
examples/tests_ci/sod_tube_sph.py (M6, CD10, periodic, t = 0.245), checked against the five constants that
script asserts (rtol 1e-11).  Needs a GPU.    python examples/run_sod_tube_b200.py [--fp fast]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shamrock_b200 import pyshamrock as shamrock  # noqa: E402

EXPECT = {"rho": 0.00016154918188486815, "vx": 0.001162704743480841, "vy": 2.988130616021184e-05,
          "vz": 1.7413547093230376e-07, "P": 0.00012483646129766217}


def main(fp_mode="strict", t_target=0.245):
    gamma, resol = 1.4, 128
    rho_l, rho_r, P_l, P_r = 1.0, 0.125, 1.0, 0.1
    spacing_ratio = (rho_l / rho_r) ** (1.0 / 3.0)

    ctx = shamrock.Context()
    ctx.pdata_layout_new()
    model = shamrock.get_Model_SPH(context=ctx, vector_type="f64_3", sph_kernel="M6", fp_mode=fp_mode)
    cfg = model.gen_default_config()
    cfg.set_artif_viscosity_VaryingCD10(alpha_min=0.0, alpha_max=1, sigma_decay=0.1, alpha_u=1, beta_AV=2)
    cfg.set_boundary_periodic()
    cfg.set_eos_adiabatic(gamma)
    model.set_solver_config(cfg)
    model.init_scheduler(int(1e8), 1)

    xs, ys, zs = model.get_box_dim_fcc_3d(1, resol, 24, 24)
    dr = 1 / xs
    xs, ys, zs = model.get_box_dim_fcc_3d(dr, resol, 24, 24)
    lo, mid_lo, mid_hi, hi = (-xs, -ys / 2, -zs / 2), (0, -ys / 2, -zs / 2), (0, ys / 2, zs / 2), (xs, ys / 2, zs / 2)
    model.resize_simulation_box(lo, hi)

    setup = model.get_setup()
    left = setup.make_generator_lattice_hcp(dr, lo, mid_hi)
    right = setup.make_generator_lattice_hcp(dr * spacing_ratio, mid_lo, hi)
    setup.apply_setup(setup.make_combiner_add(left, right))
    model.set_value_in_a_box("uint", "f64", P_l / ((gamma - 1) * rho_l), lo, mid_hi)
    model.set_value_in_a_box("uint", "f64", P_r / ((gamma - 1) * rho_r), mid_lo, hi)

    vol_half = xs * ys * zs
    model.set_particle_mass(model.total_mass_to_part_mass(rho_r * vol_half + rho_l * vol_half))
    model.set_cfl_cour(0.1)
    model.set_cfl_force(0.1)

    n_iter = model.evolve_until(t_target)

    sod = shamrock.phys.SodTube(gamma=gamma, rho_1=rho_l, P_1=P_l, rho_5=rho_r, P_5=P_r)
    rho, (vx, vy, vz), P = model.make_analysis_sodtube(sod, (1, 0, 0), t_target, 0.0, -0.5, 0.5).compute_L2_dist()
    got = {"rho": rho, "vx": vx, "vy": vy, "vz": vz, "P": P}
    rel = {k: (got[k] - EXPECT[k]) / EXPECT[k] for k in EXPECT}
    return n_iter, got, rel


if __name__ == "__main__":
    n, got, rel = main("fast" if "--fp" in sys.argv and "fast" in sys.argv else "strict")
    print(f"{n} iterations")
    for k in EXPECT:
        print(f"err_{k} = {got[k]!r}  (relative distance to the reference's constant: {rel[k]:+.2e})")
    bad = [k for k in EXPECT if abs(rel[k]) >= 1e-11]
    if bad:
        sys.exit("Test did not pass L2 margins: " + ", ".join(bad))
