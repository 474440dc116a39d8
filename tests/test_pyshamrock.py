"""The reference's Python surface on top of libshamb200 (shamrock_b200.pyshamrock).

CPU: the setup calls of the reference's Sod script give the particles of tests/sod_tube.py (the restatement
the oracle was pinned with), the periodic-box benchmark script's calls give tests/scenarios.periodic_box.
GPU: examples/run_sod_tube_b200.py — the reference's CI case through its own API — meets the reference's
constants at its own tolerance."""
import math
import os
import sys

import numpy as np
import pytest

from shamrock_b200 import pyshamrock as shamrock
from tests import scenarios as S
from tests import sod_tube as sod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sod_setup_through_the_api():
    gamma = 1.4
    ctx = shamrock.Context()
    ctx.pdata_layout_new()
    model = shamrock.get_Model_SPH(context=ctx, vector_type="f64_3", sph_kernel="M6")
    cfg = model.gen_default_config()
    cfg.set_artif_viscosity_VaryingCD10(alpha_min=0.0, alpha_max=1, sigma_decay=0.1, alpha_u=1, beta_AV=2)
    cfg.set_boundary_periodic()
    cfg.set_eos_adiabatic(gamma)
    model.set_solver_config(cfg)
    model.init_scheduler(int(1e8), 1)
    xs, ys, zs = model.get_box_dim_fcc_3d(1, 128, 24, 24)
    dr = 1 / xs
    xs, ys, zs = model.get_box_dim_fcc_3d(dr, 128, 24, 24)
    model.resize_simulation_box((-xs, -ys / 2, -zs / 2), (xs, ys / 2, zs / 2))
    setup = model.get_setup()
    g1 = setup.make_generator_lattice_hcp(dr, (-xs, -ys / 2, -zs / 2), (0, ys / 2, zs / 2))
    g2 = setup.make_generator_lattice_hcp(dr * 2.0, (0, -ys / 2, -zs / 2), (xs, ys / 2, zs / 2))
    setup.apply_setup(setup.make_combiner_add(g1, g2))
    model.set_value_in_a_box("uint", "f64", 1 / (0.4 * 1), (-xs, -ys / 2, -zs / 2), (0, ys / 2, zs / 2))
    model.set_value_in_a_box("uint", "f64", 0.1 / (0.4 * 0.125), (0, -ys / 2, -zs / 2), (xs, ys / 2, zs / 2))
    vol_b = xs * ys * zs
    pmass = model.total_mass_to_part_mass(0.125 * vol_b + 1 * vol_b)
    sc = sod.scenario()
    d = ctx.collect_data()
    assert np.array_equal(d["xyz"], sc["xyz"]) and np.array_equal(d["hpart"], sc["hpart"])
    assert np.allclose(d["uint"], sc["uint"], rtol=1e-15) and pmass == sc["cfg"]["gpart_mass"]
    assert model.get_hfact() == 1.0 and model.get_total_part_count() == 82944


def test_benchmark_setup_through_the_api():
    """sph_homogeneous_benchmark.py's calls: ideal HCP box, lattice, kernel injection, sums"""
    n_target = 20000
    sc = S.periodic_box(n_target, "M4", "cd10")
    dr = sc["dr"]
    ctx = shamrock.Context()
    model = shamrock.get_Model_SPH(context=ctx, vector_type="f64_3", sph_kernel="M4")
    bmin, bmax = shamrock.math.get_ideal_hcp_box(dr, (-0.6, -0.6, -0.6), (0.6, 0.6, 0.6))
    assert bmin == sc["bmin"] and bmax == sc["bmax"]
    model.resize_simulation_box(bmin, bmax)
    setup = model.get_setup()
    setup.apply_setup(setup.make_generator_lattice_hcp(dr, bmin, bmax))
    assert np.array_equal(ctx.collect_data()["xyz"], sc["xyz"])
    model.set_value_in_a_box("uint", "f64", 0, bmin, bmax)
    model.add_kernel_value("uint", "f64", 1.0, (0, 0, 0), 16 * dr)
    u = ctx.collect_data()["uint"]
    r = np.linalg.norm(sc["xyz"], axis=1)
    assert np.allclose(u, S.m4_w(r / (16 * dr)) / (16 * dr) ** 3, rtol=1e-13, atol=0)
    # the kernel integrates to one: sum(u) * volume per particle ~ 1
    vol = np.prod(np.array(bmax) - np.array(bmin))
    assert abs(model.get_sum("uint", "f64") * vol / len(u) - 1) < 2e-2
    xc = model.get_closest_part_to((0, 0, 0))
    assert np.linalg.norm(xc) == r.min()
    with pytest.raises(ValueError):
        shamrock.get_Model_SPH(context=ctx, vector_type="f32_3", sph_kernel="M4")


def test_sod_tube_analytic_solution_matches_restatement():
    a, b = shamrock.phys.SodTube(gamma=1.4, rho_1=1, P_1=1, rho_5=0.125, P_5=0.1), sod.SodTube(1.4, 1, 1, 0.125, 0.1)
    x = np.linspace(-0.5, 0.5, 101)
    for u, v in zip(a.get_value(0.245, x), b.get_value(0.245, x)):
        assert np.array_equal(u, v)
    ucte = shamrock.Constants(shamrock.UnitSystem(unit_time=3600 * 24 * 365, unit_length=149597870700, unit_mass=1.98847e30))
    assert abs(ucte.G() - 4 * math.pi**2) / (4 * math.pi**2) < 2e-3  # G = 4 pi^2 au^3 / (Msun yr^2)


@pytest.mark.gpu
@pytest.mark.parametrize("fp_mode", ["strict", "fast"])
def test_reference_ci_case_through_the_api(fp_mode):
    pytest.importorskip("torch")
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import run_sod_tube_b200 as ex

    n, got, rel = ex.main(fp_mode)
    assert n == 646
    for k, v in rel.items():
        assert abs(v) < 1e-11, (k, v)


def test_constants_against_the_reference_code_itself():
    """shamunits/{UnitSystem,Constants}.hpp compiled where they lie (oracle/_ref/units_ref): the constants the disc
    script reads (examples/sph/run_circular_disc_central_pot.py:40-50) in SI and in its code units (yr, au, Msun)"""
    from oracle import io_formats as io

    si = shamrock.Constants(shamrock.UnitSystem())
    for ut, ul, um in ((1.0, 1.0, 1.0), (3600 * 24 * 365, si.au(), si.sol_mass()), (7.0, 1e3, 2.5e-4)):
        ref = io.ref_units(ut, ul, um)
        if ref is None:
            pytest.skip("oracle/_ref/units_ref is not built (needs /root/reference)")
        c = shamrock.Constants(shamrock.UnitSystem(unit_time=ut, unit_length=ul, unit_mass=um))
        assert np.allclose([c.G(), c.year(), c.au(), c.sol_mass()], ref, rtol=1e-14, atol=0)
