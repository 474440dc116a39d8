"""The reference's end-to-end known answer for the SPH step: examples/tests_ci/sod_tube_sph.py.

That CI script runs a 3D Sod shock tube (M6 kernel, CD10 viscosity, periodic box, ~8e4 particles) to
t = 0.245 and requires the mean squared distances to the analytic solution to match five constants to
1e-11 relative.  It is the only golden of the reference that goes through the force loop, the CD10
operators, the leapfrog, the ghost zones and the CFL control, i.e. the whole evolve_once path.

This module restates the host side of that script (setup + analysis) in numpy; the step itself is run by
the oracle or by the CUDA path.  Nothing here reads /root/reference at run time.  Restated pieces:
  * generic::setup::generators::get_box_dim (shammodels/common/include/shammodels/common/setup/
    generators.hpp:31-51), Model::get_box_dim_fcc_3d (shammodels/sph/include/shammodels/sph/Model.hpp:90)
  * GeneratorLatticeHCP + CombinerAdd (modules/setup/GeneratorLatticeHCP.hpp:39-120): lattice points with
    lower <= r < upper, hpart = dr of the generator; apply_setup reorders each patch by Morton code
    (modules/ParticleReordering.cpp:24-52) - the order only matters for equal Morton codes, none here
  * Model::set_value_in_a_box (Model.hpp:669-705, lower <= r < upper), total_mass_to_part_mass (Model.cpp:103)
  * Solver::evolve_until (Solver.hpp:305-370): dt clipped to land on the target time
  * shamphys::SodTube (shamphys/src/SodTube.cpp:23-150) incl. newton_rhaphson returning `float`
    (shammath/include/shammath/solve.hpp:28) and derivative_upwind (derivatives.hpp:38)
  * modules::AnalysisSodTube::compute_L2_dist (shammodels/sph/src/modules/AnalysisSodTube.cpp:28-123)
"""
import math

import numpy as np

from shamrock_b200 import lattice

# examples/tests_ci/sod_tube_sph.py:131-137
EXPECTED = {"rho": 0.00016154918188486815, "vx": 0.001162704743480841, "vy": 2.988130616021184e-05,
            "vz": 1.7413547093230376e-07, "P": 0.00012483646129766217}
RTOL_REFERENCE = 1e-11
T_TARGET = 0.245
GAMMA = 1.4


def get_box_dim(r_particle, xcnt, ycnt, zcnt):
    i, j, k = xcnt, ycnt, zcnt
    r = (2 * i + ((j + k) % 2), math.sqrt(3.0) * (j + (1.0 / 3.0) * (k % 2)), 2 * math.sqrt(6.0) * k / 3)
    return tuple(c * r_particle for c in r)


def scenario(resol=128, ny=24, nz=24, kernel="M6"):
    """sod_tube_sph.py:20-100 -> scenario dict of tests/scenarios.py"""
    rho_g, rho_d = 1.0, 0.125
    fact = (rho_g / rho_d) ** (1.0 / 3.0)
    P_g, P_d = 1.0, 0.1
    u_g = P_g / ((GAMMA - 1) * rho_g)
    u_d = P_d / ((GAMMA - 1) * rho_d)
    xs, ys, zs = get_box_dim(1.0, resol, ny, nz)
    dr = 1 / xs
    xs, ys, zs = get_box_dim(dr, resol, ny, nz)
    bmin, bmax = (-xs, -ys / 2, -zs / 2), (xs, ys / 2, zs / 2)
    left = lattice.hcp_positions(dr, bmin, (0.0, ys / 2, zs / 2))
    right = lattice.hcp_positions(dr * fact, (0.0, -ys / 2, -zs / 2), bmax)
    pos = np.concatenate([left, right])
    h = np.concatenate([np.full(len(left), dr), np.full(len(right), dr * fact)])
    u = np.zeros(len(pos))
    x, y, z = pos[:, 0], pos[:, 1], pos[:, 2]
    for val, lo, hi in ((u_g, bmin, (0.0, ys / 2, zs / 2)), (u_d, (0.0, -ys / 2, -zs / 2), bmax)):
        sel = (lo[0] <= x) & (x < hi[0]) & (lo[1] <= y) & (y < hi[1]) & (lo[2] <= z) & (z < hi[2])
        u[sel] = val
    vol_b = xs * ys * zs
    totmass = (rho_d * vol_b) + (rho_g * vol_b)
    pmass = totmass / len(pos)
    kid = {"M4": 0, "M6": 1}[kernel]
    cfg = dict(kernel=kid, gpart_mass=pmass, eos=0, gamma=GAMMA, av=3, alpha_u=1.0, alpha_AV=1.0, beta_AV=2.0,
               alpha_min=0.0, alpha_max=1.0, sigma_decay=0.1, bc=1, cfl_cour=0.1, cfl_force=0.1)
    return dict(name=f"sod_ci_{kernel}_{len(pos)}", cfg=cfg, bmin=bmin, bmax=bmax, grid=(1, 1, 1), xyz=pos,
                vxyz=np.zeros_like(pos), hpart=h, uint=u, kill=[], sort_mode="bitonic", kernel=kernel, dr=dr)


def evolve_until(model, t_target, niter_max=-1, log=None):
    """Solver::evolve_until: `model` is the oracle Solver or the CUDA Model (same method names)"""
    n = 0
    st = model.state()
    while st["time"] < t_target:
        if st["time"] + st["dt"] > t_target:
            model.set_next_dt(t_target - st["time"])
        st = model.evolve_once()
        n += 1
        if log and n % 50 == 0:
            log(f"  iter {n} t = {st['time']:.6f} dt = {st['dt']:.3e}")
        if 0 <= niter_max <= n:
            break
    return n


class SodTube:
    """shamphys::SodTube"""

    def __init__(self, gamma, rho_1, P_1, rho_5, P_5):
        self.gamma, self.rho_1, self.P_1, self.rho_5, self.P_5 = gamma, rho_1, P_1, rho_5, P_5
        self.c_1 = math.sqrt(gamma * P_1 / rho_1)
        self.c_5 = math.sqrt(gamma * P_5 / rho_5)

    def solve_P_4(self):
        g, c_1, c_5, P_1, P_5 = self.gamma, self.c_1, self.c_5, self.P_1, self.P_5

        def f(P_4):
            z = P_4 / P_5 - 1.0
            gm1, gp1, g2 = g - 1.0, g + 1.0, 2.0 * g
            fact1 = gm1 / g2 * (c_5 / c_1) * z / math.sqrt(1.0 + gp1 / g2 * z)
            fact = math.pow(1.0 - fact1, g2 / gm1)
            return P_1 * fact - P_4

        def df(P_4):
            return (f(P_4 + 1e-6) - f(P_4)) / 1e-6

        xk, eps = P_1, 100000.0
        while eps > 1e-6:
            xkp1 = xk - (f(xk) / df(xk))
            eps = abs(xk - xkp1)
            xk = xkp1
        return float(np.float32(xk))  # newton_rhaphson returns `float` (solve.hpp:28)

    def get_value(self, t, x):
        """vectorised over x; returns rho, vx, P"""
        g = self.gamma
        P_4 = self.solve_P_4()
        z = P_4 / self.P_5 - 1.0
        gm1, gp1 = g - 1.0, g + 1.0
        gmfact1, gmfact2 = 0.5 * gm1 / g, 0.5 * gp1 / g
        fact = math.sqrt(1.0 + gmfact2 * z)
        vx_4 = self.c_5 * z / (g * fact)
        rho_4 = self.rho_5 * (1.0 + gmfact2 * z) / (1.0 + gmfact1 * z)
        w = self.c_5 * fact
        P_3, vx_3 = P_4, vx_4
        rho_3 = self.rho_1 * math.pow(P_3 / self.P_1, 1.0 / g)
        c3 = math.sqrt(g * P_3 / rho_3)
        xsh, xcd, xft, xhd = w * t, vx_3 * t, (vx_3 - c3) * t, -self.c_1 * t
        x = np.asarray(x, dtype=np.float64)
        vx_r = 2.0 / gp1 * (self.c_1 + x / t)
        locfact = 1.0 - 0.5 * gm1 * vx_r / self.c_1
        with np.errstate(invalid="ignore"):
            rho_r = self.rho_1 * np.power(locfact, 2.0 / gm1)
            p_r = self.P_1 * np.power(locfact, 2.0 * g / gm1)
        conds = [x < xhd, x < xft, x < xcd, x < xsh]
        rho = np.select(conds, [self.rho_1, rho_r, rho_3, rho_4], self.rho_5)
        p = np.select(conds, [self.P_1, p_r, P_3, P_4], self.P_5)
        vx = np.select(conds, [0.0, vx_r, vx_3, vx_4], 0.0)
        return rho, vx, p


def compute_L2_dist(xyz, vxyz, hpart, uint, pmass, hfact, gamma=GAMMA, t=T_TARGET, x_ref=0.0, x_min=-0.5, x_max=0.5):
    """AnalysisSodTube::compute_L2_dist with direction (1, 0, 0): mean squared distances (no square root)"""
    sod = SodTube(gamma, 1.0, 1.0, 0.125, 0.1)
    q = hfact / hpart
    rho = pmass * q * q * q
    P = (gamma - 1) * rho * uint
    x = xyz[:, 0] - x_ref
    sel = ((x + x_ref) > x_min) & ((x + x_ref) < x_max)
    r_rho, r_vx, r_P = sod.get_value(t, x[sel])
    d_rho, d_P = rho[sel] - r_rho, P[sel] - r_P
    dv = vxyz[sel].copy()
    dv[:, 0] -= r_vx
    n = float(sel.sum())
    # the reference accumulates in particle order on the host; fsum (exactly rounded) is within 1 ulp of it
    out = {"rho": math.fsum(d_rho * d_rho) / n, "P": math.fsum(d_P * d_P) / n,
           "vx": math.fsum(dv[:, 0] ** 2) / n, "vy": math.fsum(dv[:, 1] ** 2) / n, "vz": math.fsum(dv[:, 2] ** 2) / n}
    return out


def relative_errors(l2):
    return {k: (l2[k] - EXPECTED[k]) / EXPECTED[k] for k in EXPECTED}


def fingerprint(model):
    """bit-level fingerprint of the solver state (hex floats)"""
    st = model.state()
    out = {"time": float(st["time"]).hex(), "dt": float(st["dt"]).hex()}
    for nm in ("xyz", "vxyz", "hpart", "uint", "axyz", "duint", "alpha_AV"):
        a = model.get(0, nm)
        out[nm] = math.fsum(a.reshape(-1).tolist()).hex()  # exactly rounded sum: independent of the order
        out[nm + "_absmax"] = float(np.abs(a).max()).hex()
    return out
