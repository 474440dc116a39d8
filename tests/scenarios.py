"""Seeded synthetic initial conditions shared by the parity tests, smoke() and bench.py.

Each scenario returns a dict: cfg (solver configuration, the key names of
shammodels::sph::SolverConfig used by both the oracle and the CUDA model), box, patch grid,
particle arrays (numpy, host) and optional kill spheres.  Geometry follows the reference scripts:
examples/benchmarks/sph_homogeneous_benchmark.py:65-160 (periodic HCP box + Sedov-like injection),
examples/tests_ci/sod_tube_sph.py:20-110 (Sod tube), examples/sph/run_circular_disc_central_pot.py
(disc).  Nothing here reads /root/reference at run time.
"""
import math

import numpy as np

from shamrock_b200 import lattice

KERNEL_ID = {"M4": 0, "M6": 1}
HFACT = {"M4": 1.2, "M6": 1.0}
RKERN = {"M4": 2.0, "M6": 3.0}


def m4_w(q):
    """M4 kernel shape (for add_kernel_value-like injection); shammath/sphkernels.hpp:42-56"""
    t1 = np.clip(2 - q, 0, None) ** 3 / 4
    t2 = -np.clip(1 - q, 0, None) ** 3
    return (t1 + t2) / math.pi


def periodic_box_geometry(n_target, stretch=(1, 1, 1)):
    """lattice spacing and ideal periodic HCP box of periodic_box (no particles generated)"""
    half = np.array([0.6 * stretch[0], 0.6 * stretch[1], 0.6 * stretch[2]])
    vol = float(np.prod(2 * half))
    dr = (vol / (n_target * 4 * math.sqrt(2))) ** (1.0 / 3.0)
    bmin, bmax = lattice.get_ideal_hcp_box(dr, tuple(-half), tuple(half))
    return dr, bmin, bmax


def periodic_box_on_device(m, n_target, kernel="M4", stretch=(1, 1, 1)):
    """the unjittered periodic_box scenario generated on the device (shamb200_model_add_lattice_hcp and the
    device-side setters): every rank builds the particles of its own patches.  Same particles, fields and
    particle mass as periodic_box(jitter=0) up to the last bit of the injected energy (numpy's pow vs x*x*x)."""
    dr, bmin, bmax = periodic_box_geometry(n_target, stretch)
    n = m.add_lattice_hcp(dr, bmin, bmax)
    vol = float(np.prod(np.array(bmax) - np.array(bmin)))
    pmass = 1.0 * vol / n
    m.set_particle_mass(pmass)
    big = ([-1e300] * 3, [1e300] * 3)
    m.set_value_in_a_box("hpart", HFACT[kernel] * (pmass / 1.0) ** (1.0 / 3.0), *big)
    m.set_value_in_a_box("uint", 1.0, *big)
    m.add_kernel_value("uint", 50.0 * pmass, (0.0, 0.0, 0.0), 16 * dr)
    return dict(n=n, dr=dr, bmin=bmin, bmax=bmax, pmass=pmass)


def periodic_box(n_target, kernel="M4", av="cd10", jitter=0.0, seed=42, grid=(1, 1, 1), stretch=(1, 1, 1),
                 two_stage=True, inject=True, sort_mode="bitonic", local_boxes=None, count_reduce=None):
    """sph_homogeneous_benchmark.py: HCP lattice in a periodic box, adiabatic gamma=5/3, CD10 AV,
    uint kernel injection at the origin, C_cour=C_force=0.1.

    Multi-rank setup: `local_boxes(bmin, bmax)` returns the [lo, hi) boxes of this rank's patches and only
    their particles are generated; `count_reduce(n_local)` returns the global particle count (an
    all-reduce) that fixes the particle mass.  Same particles as the single-process call, rank by rank."""
    # HCP: one particle per dr^3 * sqrt(2) * 4  (cell volume per particle = 4 sqrt(2) dr^3)
    dr, bmin, bmax = periodic_box_geometry(n_target, stretch)
    if local_boxes == "device":  # geometry and configuration only: the particles are generated on the device
        pos = np.zeros((0, 3))
        n = 1
    elif local_boxes is None:
        pos = lattice.hcp_positions(dr, bmin, bmax)
        n = len(pos)
    else:
        parts = []
        for lo, hi in local_boxes(bmin, bmax):
            # the lattice points of the whole box that fall into [lo, hi): clip the patch to the box first
            lo_c = [max(a, b) for a, b in zip(lo, bmin)]
            hi_c = [min(a, b) for a, b in zip(hi, bmax)]
            parts.append(lattice.hcp_positions(dr, lo_c, hi_c))
        pos = np.concatenate(parts) if parts else np.zeros((0, 3))
        n = int(count_reduce(len(pos)))
    rng = np.random.default_rng(seed)
    if jitter > 0:
        pos = pos + rng.uniform(-jitter * dr, jitter * dr, size=pos.shape)
        for c in range(3):  # keep the particles inside [bmin, bmax)
            L = bmax[c] - bmin[c]
            pos[:, c] = bmin[c] + np.mod(pos[:, c] - bmin[c], L)
            pos[pos[:, c] >= bmax[c], c] = bmin[c]
    vol = float(np.prod(np.array(bmax) - np.array(bmin)))
    rho = 1.0
    pmass = rho * vol / n
    h = np.full(len(pos), HFACT[kernel] * (pmass / rho) ** (1.0 / 3.0))
    u = np.full(len(pos), 1.0)
    if inject:
        r = np.linalg.norm(pos, axis=1)
        hi = 16 * dr
        u = u + 1.0 * m4_w(r / hi) / hi**3 * pmass * 50.0
    v = np.zeros_like(pos)
    if jitter > 0:
        v = rng.normal(0, 0.05, size=pos.shape)
    avid = {"constant": 1, "mm97": 2, "cd10": 3}[av]
    cfg = dict(kernel=KERNEL_ID[kernel], gpart_mass=pmass, eos=0, gamma=5.0 / 3.0, av=avid, alpha_u=1.0,
               alpha_AV=1.0, beta_AV=2.0, alpha_min=0.0, alpha_max=1.0, sigma_decay=0.1, bc=1, cfl_cour=0.1,
               cfl_force=0.1, use_two_stage_search=int(two_stage))
    return dict(name=f"periodic_{kernel}_{av}_{n}", cfg=cfg, bmin=bmin, bmax=bmax, grid=grid, xyz=pos, vxyz=v,
                hpart=h, uint=u, kill=[], sort_mode=sort_mode, kernel=kernel, dr=dr)


def sod_tube(resol=24, kernel="M6", grid=(2, 1, 1), sort_mode="bitonic"):
    """sod_tube_sph.py geometry (smaller): two HCP lattices, rho 1 / 0.125, P 1 / 0.1, gamma 1.4,
    CD10, periodic."""
    gamma = 1.4
    rho_g, rho_d = 1.0, 0.125
    P_g, P_d = 1.0, 0.1
    fact = (rho_g / rho_d) ** (1.0 / 3.0)
    dr = 1.0 / resol
    (xs, ys, zs) = (1.0, 6 * dr * 2, 6 * dr * 2)
    bmin, bmax = lattice.get_ideal_hcp_box(dr, (-xs, -ys / 2, -zs / 2), (xs, ys / 2, zs / 2))
    xs = bmax[0]
    left = lattice.hcp_positions(dr, (bmin[0], bmin[1], bmin[2]), (0.0, bmax[1], bmax[2]))
    right = lattice.hcp_positions(dr * fact, (0.0, bmin[1], bmin[2]), (bmax[0], bmax[1], bmax[2]))
    pos = np.concatenate([left, right])
    n = len(pos)
    vol_l = (0.0 - bmin[0]) * (bmax[1] - bmin[1]) * (bmax[2] - bmin[2])
    vol_r = (bmax[0] - 0.0) * (bmax[1] - bmin[1]) * (bmax[2] - bmin[2])
    pmass = (rho_g * vol_l + rho_d * vol_r) / n
    h = np.concatenate([np.full(len(left), dr), np.full(len(right), dr * fact)])
    u = np.concatenate([np.full(len(left), P_g / ((gamma - 1) * rho_g)),
                        np.full(len(right), P_d / ((gamma - 1) * rho_d))])
    cfg = dict(kernel=KERNEL_ID[kernel], gpart_mass=pmass, eos=0, gamma=gamma, av=3, alpha_u=1.0, alpha_AV=1.0,
               beta_AV=2.0, alpha_min=0.0, alpha_max=1.0, sigma_decay=0.1, bc=1, cfl_cour=0.3, cfl_force=0.25)
    return dict(name=f"sod_{kernel}_{n}", cfg=cfg, bmin=bmin, bmax=bmax, grid=grid, xyz=pos,
                vxyz=np.zeros_like(pos), hpart=h, uint=u, kill=[], sort_mode=sort_mode, kernel=kernel, dr=dr)


def disc(n=4000, kernel="M4", seed=7, grid=(1, 1, 1), sort_mode="bitonic", regular=False):
    """Protoplanetary-disc-like cloud around a central point mass: free boundaries, LP07 locally
    isothermal EOS, ConstantDisc AV, accretion radius and a kill sphere
    (run_circular_disc_central_pot.py:190-248, Monte-Carlo positions with a fixed seed).
    regular=True: the same flared volume filled with an HCP lattice instead of Monte-Carlo points (large
    runs: Poisson clumps of an unrelaxed random sample leave a few particles whose h never converges —
    in the reference's algorithm as well, the oracle shows it at 2e5 particles)."""
    rng = np.random.default_rng(seed)
    rin, rout, H_r = 1.0, 3.0, 0.08
    if regular:
        vol = math.pi * (rout**2 - rin**2) * 3 * H_r * 2.0  # mean radius 2: thickness 3 H_r r
        dr = (vol / (n * 4 * math.sqrt(2))) ** (1.0 / 3.0)
        zmax = 1.5 * H_r * rout
        p = lattice.hcp_positions(dr, (-rout, -rout, -zmax), (rout, rout, zmax))
        rr = np.hypot(p[:, 0], p[:, 1])
        p = p[(rr > rin * 1.05) & (rr < rout) & (np.abs(p[:, 2]) < 1.5 * H_r * rr)]
        n = len(p)
        r, phi, z = np.hypot(p[:, 0], p[:, 1]), np.arctan2(p[:, 1], p[:, 0]), p[:, 2]
    else:
        r = np.sqrt(rng.uniform(rin**2, rout**2, n))
        phi = rng.uniform(0, 2 * math.pi, n)
        z = rng.uniform(-1.5, 1.5, n) * H_r * r  # truncated: isolated particles never converge in h
    pos = np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)
    G, Mc = 1.0, 1.0
    vk = np.sqrt(G * Mc / r)
    v = np.stack([-vk * np.sin(phi), vk * np.cos(phi), np.zeros(n)], axis=1)
    disc_mass = 0.01
    pmass = disc_mass / n
    vol = math.pi * (rout**2 - rin**2) * 3 * H_r * 2.0
    nd = n / (math.pi * (rout**2 - rin**2)) / (3 * H_r * r)  # local number density
    h = HFACT[kernel] * nd ** (-1.0 / 3.0)
    if regular:
        h = np.full(n, HFACT[kernel] * (4 * math.sqrt(2)) ** (1.0 / 3.0) * dr)
    u = np.full(n, 1e-3)
    cfg = dict(kernel=KERNEL_ID[kernel], gpart_mass=pmass, eos=2, cs0=0.05, eos_q=0.25, eos_r0=1.0, av=4,
               alpha_u=1.0, alpha_AV=1.0, beta_AV=2.0, bc=0, cfl_cour=0.3, cfl_force=0.25, has_point_mass=1,
               pm_mass=Mc, pm_racc=1.02 * rin, constant_G=G)
    b = rout * 1.5
    return dict(name=f"disc_{kernel}_{n}", cfg=cfg, bmin=(-b, -b, -b), bmax=(b, b, b), grid=grid, xyz=pos, vxyz=v,
                hpart=h, uint=u, kill=[((0.0, 0.0, 0.0), 2.9)], sort_mode=sort_mode, kernel=kernel,
                dr=(vol / n) ** (1.0 / 3.0))


# ---- builders -------------------------------------------------------------------------------------
def make_oracle(sc):
    from oracle import pyoracle as po

    s = po.Solver(sc["cfg"], sc["bmin"], sc["bmax"], sc["grid"])
    for c, r in sc["kill"]:
        s.add_kill_sphere(c, r)
    s.push_particles(sc["xyz"], sc["vxyz"], sc["hpart"], sc["uint"])
    return s


def patch_loads(sc, world=1):
    """particles of the scenario per patch of its grid (patch id order) — the load values of the balancer"""
    from shamrock_b200 import _capi

    boxes, _ = _capi.plan_patch_grid(sc["bmin"], sc["bmax"], sc["grid"], world)
    x = np.asarray(sc["xyz"])
    return np.array([int(np.all((x >= np.array(lo)) & (x < np.array(hi)), axis=1).sum()) for lo, hi in boxes],
                    dtype=np.uint64)


def make_cuda(sc, ctx=None, keep_step_data=True, rank=0, world=1, nccl_id=None, fp_mode="strict", balance=False):
    """balance=True: patches dealt to the ranks by the Hilbert-curve load balancer on the particle counts of the
    setup (shamb200_plan_load_balance) instead of contiguous blocks of patch ids"""
    from shamrock_b200 import _capi

    ctx = ctx or _capi.Context(0)
    cfg = _capi.default_config()
    for k, v in sc["cfg"].items():
        cur = getattr(cfg, k)
        setattr(cfg, k, int(v) if isinstance(cur, int) else float(v))
    cfg.sort_mode = _capi.SORT_MODES[sc.get("sort_mode", "bitonic")]
    cfg.keep_step_data = int(keep_step_data)
    cfg.fp_mode = _capi.FP_MODES[fp_mode]
    for i, (c, r) in enumerate(sc["kill"]):
        for d in range(3):
            cfg.kill_center[i][d] = c[d]
        cfg.kill_radius[i] = r
    cfg.n_kill_spheres = len(sc["kill"])
    m = _capi.Model(ctx, cfg)
    if world > 1:
        m.init_comm(rank, world, nccl_id)
    m.set_box(sc["bmin"], sc["bmax"], sc["grid"])
    if balance:
        owner, _ = _capi.plan_load_balance(m.patch_coords(), patch_loads(sc, world), world)
        m.set_patch_owners(owner)
    if len(sc["xyz"]):
        m.push_particles(sc["xyz"], sc["vxyz"], sc["hpart"], sc["uint"])
    return m
