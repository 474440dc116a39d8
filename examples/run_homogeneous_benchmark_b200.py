#!/usr/bin/env python
"""The reference's full-step benchmark protocol (examples/benchmarks/sph_homogeneous_benchmark.py) through
its Python surface on the B200 backend: HCP lattice in a periodic box, M4, CD10, one warm-up timestep, then
ten replays with set_next_dt(0); prints the best rate.  `bench.py` measures the same workload through the
C ABI (and adds the roofline / e2e legs).    python examples/run_homogeneous_benchmark_b200.py [N_target]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shamrock_b200 import pyshamrock as shamrock  # noqa: E402


def main(n_target=4_000_000, fp_mode="fast", replays=10):
    gamma, rho_g = 5.0 / 3.0, 1.0
    half = 0.6
    part_vol = (2 * half) ** 3 / n_target
    dr = (0.74 * part_vol / ((4.0 / 3.0) * 3.1416)) ** (1.0 / 3.0)  # the reference script's spacing rule

    ctx = shamrock.Context()
    ctx.pdata_layout_new()
    model = shamrock.get_Model_SPH(context=ctx, vector_type="f64_3", sph_kernel="M4", fp_mode=fp_mode, sort_mode="radix")
    cfg = model.gen_default_config()
    cfg.set_artif_viscosity_VaryingCD10(alpha_min=0.0, alpha_max=1, sigma_decay=0.1, alpha_u=1, beta_AV=2)
    cfg.set_boundary_periodic()
    cfg.set_eos_adiabatic(gamma)
    model.set_solver_config(cfg)
    model.init_scheduler(int(2e7), 1)

    bmin, bmax = shamrock.math.get_ideal_hcp_box(dr, (-half,) * 3, (half,) * 3)
    model.resize_simulation_box(bmin, bmax)
    setup = model.get_setup()
    setup.apply_setup(setup.make_generator_lattice_hcp(dr, bmin, bmax))

    vol_b = (bmax[0] - bmin[0]) * (bmax[1] - bmin[1]) * (bmax[2] - bmin[2])
    pmass = model.total_mass_to_part_mass(rho_g * vol_b)
    model.set_value_in_a_box("uint", "f64", 0, bmin, bmax)
    model.add_kernel_value("uint", "f64", 1, (0, 0, 0), 16 * dr)
    tot_u = pmass * model.get_sum("uint", "f64")
    model.set_particle_mass(pmass)
    model.set_cfl_cour(0.1)
    model.set_cfl_force(0.1)

    model.timestep()  # converges the smoothing lengths and computes the first dt
    rates, counts = [], []
    for _ in range(replays):
        model.set_next_dt(0.0)  # replay the same step
        model.timestep()
        rates.append(model.solver_logs_last_rate())
        counts.append(model.solver_logs_last_obj_count())
    return {"npart": counts[-1], "total_u": tot_u, "best_rate": max(rates), "rates": rates}


if __name__ == "__main__":
    res = main(int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000)
    print(f"N = {res['npart']}, total u = {res['total_u']:.6g}")
    print("rates (particles / s):", ", ".join(f"{r:.4g}" for r in res["rates"]))
    print(f"best rate: {res['best_rate']:.4g} particles / s")
