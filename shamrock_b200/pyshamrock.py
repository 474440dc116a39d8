"""The part of the reference's Python surface (`import shamrock`) that its SPH scripts use to reach
`Solver::evolve_once`, on top of libshamb200 — same names, keyword arguments and meaning, so that
examples/tests_ci/sod_tube_sph.py, examples/benchmarks/sph_homogeneous_benchmark.py and
examples/sph/run_circular_disc_central_pot.py read the same with `from shamrock_b200 import pyshamrock as shamrock`.

Mirrors (paths relative to /root/reference/src): shammodels/sph/src/pySPHModel.cpp (Model / SolverConfig /
SPHSetup bindings), shammodels/sph/include/shammodels/sph/Model.hpp (:82-135 scheduler / box / cfl setters,
:669-705 set_value_in_a_box, :740-785 add_kernel_value / get_sum, :997-1009 timestep / evolve_once),
shammodels/sph/include/shammodels/sph/Solver.hpp:305-370 (evolve_until), shammodels/sph/src/Model.cpp:103
(total_mass_to_part_mass), modules/setup/GeneratorLatticeHCP.hpp and CombinerAdd (setup nodes),
shammath/include/shammath/crystalLattice.hpp (HCP lattice), shamphys/src/SodTube.cpp and
modules/AnalysisSodTube.cpp (analysis).

Setup is host work (numpy), as in the reference; the patch data moves to the GPU at the first timestep and
lives there.  One process drives one GPU; there is no CPU fallback (the step needs libshamb200 and a device).
Phantom dumps (`load_phantom_dump`, `PhantomDump`, `compare_phantom_dumps`, `model.make_phantom_dump`,
`gen_config_from_phantom_dump`, `init_from_phantom_dump`: shammodels/sph/src/pyPhantomDump.cpp:24-58,
pySPHModel.cpp:1352-1379) and `model.do_vtk_dump` go through libshamb200's csrc/io_formats.cu.
Not mirrored: plots, sinks, the solver-graph introspection (`get_solver_tex` / `get_solver_dot_graph`
return a short placeholder so that scripts printing them keep running).
"""
import math as _math
import types as _types

import numpy as np

from . import _capi, lattice

_KERNELS = {"M4": (0, 1.2, 2.0), "M6": (1, 1.0, 3.0)}  # id, hfact, Rkern


# ---- shamrock.math / shamrock.sys / shamrock.phys ------------------------------------------------------
def _get_ideal_hcp_box(dr, box_min, box_max):
    return lattice.get_ideal_hcp_box(dr, tuple(box_min), tuple(box_max))


math = _types.SimpleNamespace(get_ideal_hcp_box=_get_ideal_hcp_box)
sys = _types.SimpleNamespace(world_rank=lambda: 0, world_size=lambda: 1, mpi_barrier=lambda: None)


class SodTube:
    """shamrock.phys.SodTube (shamphys/src/SodTube.cpp:23-150)"""

    def __init__(self, gamma, rho_1, P_1, rho_5, P_5):
        if P_5 > P_1:
            raise ValueError("not correct")
        self.gamma, self.rho_1, self.P_1, self.rho_5, self.P_5 = gamma, rho_1, P_1, rho_5, P_5
        self.c_1 = _math.sqrt(gamma * P_1 / rho_1)
        self.c_5 = _math.sqrt(gamma * P_5 / rho_5)

    def _solve_P_4(self):
        g, c_1, c_5, P_1, P_5 = self.gamma, self.c_1, self.c_5, self.P_1, self.P_5

        def f(P_4):
            z = P_4 / P_5 - 1.0
            gm1, gp1, g2 = g - 1.0, g + 1.0, 2.0 * g
            fact1 = gm1 / g2 * (c_5 / c_1) * z / _math.sqrt(1.0 + gp1 / g2 * z)
            return P_1 * _math.pow(1.0 - fact1, g2 / gm1) - P_4

        xk, eps = P_1, 100000.0
        while eps > 1e-6:  # shammath::newton_rhaphson with derivative_upwind(dx = 1e-6)
            xkp1 = xk - (f(xk) / ((f(xk + 1e-6) - f(xk)) / 1e-6))
            eps, xk = abs(xk - xkp1), xkp1
        return float(np.float32(xk))  # the reference's solver returns `float` (shammath/solve.hpp:28)

    def get_value(self, t, x):
        """(rho, vx, P) at time t and position(s) x"""
        g = self.gamma
        P_4 = self._solve_P_4()
        z = P_4 / self.P_5 - 1.0
        gm1, gp1 = g - 1.0, g + 1.0
        gmfact1, gmfact2 = 0.5 * gm1 / g, 0.5 * gp1 / g
        fact = _math.sqrt(1.0 + gmfact2 * z)
        vx_4 = self.c_5 * z / (g * fact)
        rho_4 = self.rho_5 * (1.0 + gmfact2 * z) / (1.0 + gmfact1 * z)
        w = self.c_5 * fact
        P_3, vx_3 = P_4, vx_4
        rho_3 = self.rho_1 * _math.pow(P_3 / self.P_1, 1.0 / g)
        c3 = _math.sqrt(g * P_3 / rho_3)
        xsh, xcd, xft, xhd = w * t, vx_3 * t, (vx_3 - c3) * t, -self.c_1 * t
        x = np.asarray(x, dtype=np.float64)
        vx_r = 2.0 / gp1 * (self.c_1 + x / t)
        locfact = 1.0 - 0.5 * gm1 * vx_r / self.c_1
        with np.errstate(invalid="ignore"):
            rho_r = self.rho_1 * np.power(locfact, 2.0 / gm1)
            p_r = self.P_1 * np.power(locfact, 2.0 * g / gm1)
        conds = [x < xhd, x < xft, x < xcd, x < xsh]
        return (np.select(conds, [self.rho_1, rho_r, rho_3, rho_4], self.rho_5),
                np.select(conds, [0.0, vx_r, vx_3, vx_4], 0.0),
                np.select(conds, [self.P_1, p_r, P_3, P_4], self.P_5))


phys = _types.SimpleNamespace(SodTube=SodTube)


# ---- shamrock.backends / shamrock.math.AABB / shamrock.tree (the tree micro-benchmark surface) ----------
class DeviceBuffer_f64_3:
    """sham::DeviceBuffer<f64_3> as the reference's Python sees it (shampylib/src/pyShambackends.cpp: resize,
    get_size, copy_from_stdvec, copy_to_stdvec): 3 doubles per element, packed, in device memory."""

    def __init__(self):
        self._t = None

    def resize(self, n):
        import torch

        old = self._t
        self._t = torch.zeros((int(n), 3), dtype=torch.float64, device="cuda")
        if old is not None:
            k = min(len(old), int(n))
            self._t[:k] = old[:k]

    def get_size(self):
        return 0 if self._t is None else int(self._t.shape[0])

    def copy_from_stdvec(self, values):
        import torch

        a = np.ascontiguousarray(values, dtype=np.float64).reshape(-1, 3)
        if a.shape[0] != self.get_size():
            raise ValueError("buffer size mismatch")
        self._t.copy_(torch.from_numpy(a))

    def copy_to_stdvec(self):
        return [tuple(r) for r in self._t.cpu().numpy()]


class AABB_f64_3:
    """shammath::AABB<f64_3> (shampylib/src/pyShammath.cpp): lower, upper"""

    def __init__(self, lower, upper):
        self.lower, self.upper = tuple(float(v) for v in lower), tuple(float(v) for v in upper)


class CLBVH_u32_f64_3:
    """shamtree::CompressedLeafBVH<u32, f64_3, 3> — the instantiation the SPH solver uses
    (SolverConfig.hpp:443) — with the Python surface of shampylib/src/pyShamtree.cpp:28-60
    (rebuild_from_positions, get_leaf_cell_count, get_internal_cell_count, get_total_cell_count) on top of
    shamb200_tree_build.  `positions`: a DeviceBuffer_f64_3, a CUDA torch tensor (n, 3) f64, or host
    coordinates (copied to the device first).  sort_mode "bitonic" reproduces the reference's order of
    equal Morton codes; "radix" is the fast stable sort (same tree)."""

    def __init__(self, device=0, sort_mode="bitonic"):
        self._ctx = _capi.Context(device)
        self._sort_mode = sort_mode
        self._tv = None
        self._keep = None

    def rebuild_from_positions(self, positions, bounding_box, compression_level):
        import torch

        if isinstance(positions, DeviceBuffer_f64_3):
            t = positions._t
        elif isinstance(positions, torch.Tensor):
            t = positions
        else:
            t = torch.from_numpy(np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)).cuda()
        if t is None or t.dtype != torch.float64 or not t.is_cuda or t.dim() != 2 or t.shape[1] != 3:
            raise ValueError("positions must be n x 3 float64 in device memory")
        t = t.contiguous()
        torch.cuda.synchronize()
        self._keep = t
        self._tv = self._ctx.tree_build(t, int(t.shape[0]), bounding_box.lower, bounding_box.upper,
                                        reduction_level=int(compression_level), sort_mode=self._sort_mode)
        self._ctx.synchronize()

    def _need(self):
        if self._tv is None:
            raise RuntimeError("the tree is empty (rebuild_from_positions has not been called)")
        return self._tv

    def get_leaf_cell_count(self):
        return int(self._need().leaf_count)

    def get_internal_cell_count(self):
        return int(self._need().int_count)

    def get_total_cell_count(self):
        tv = self._need()
        return int(tv.leaf_count + tv.int_count)


backends = _types.SimpleNamespace(DeviceBuffer_f64_3=DeviceBuffer_f64_3)
math.AABB_f64_3 = AABB_f64_3
tree = _types.SimpleNamespace(CLBVH_u32_f64_3=CLBVH_u32_f64_3)


# ---- shamrock.Context ------------------------------------------------------------------------------------
class Context:
    """shamrock.Context: holds the scheduler / patch data of one model"""

    def __init__(self):
        self._model = None

    def pdata_layout_new(self):
        pass

    def collect_data(self):
        """dict field name -> numpy array of every particle (ctx.collect_data())"""
        return self._model._collect() if self._model else {}


# ---- solver configuration --------------------------------------------------------------------------------
class SolverConfig:
    """model.gen_default_config(): the setters of pySPHModel.cpp used by the scripts of the five configs"""

    def __init__(self, kernel):
        self._kernel = kernel
        self._c = dict(eos=0, gamma=5.0 / 3.0, av=1, alpha_u=1.0, alpha_AV=1.0, beta_AV=2.0, bc=0)
        self._kill = []
        self._units_G = 1.0

    def set_artif_viscosity_None(self):
        self._c.update(av=0)

    def set_artif_viscosity_Constant(self, alpha_u, alpha_AV, beta_AV):
        self._c.update(av=1, alpha_u=alpha_u, alpha_AV=alpha_AV, beta_AV=beta_AV)

    def set_artif_viscosity_VaryingMM97(self, alpha_min, alpha_max, sigma_decay, alpha_u, beta_AV):
        self._c.update(av=2, alpha_min=alpha_min, alpha_max=alpha_max, sigma_decay=sigma_decay, alpha_u=alpha_u,
                       beta_AV=beta_AV)

    def set_artif_viscosity_VaryingCD10(self, alpha_min, alpha_max, sigma_decay, alpha_u, beta_AV):
        self._c.update(av=3, alpha_min=alpha_min, alpha_max=alpha_max, sigma_decay=sigma_decay, alpha_u=alpha_u,
                       beta_AV=beta_AV)

    def set_artif_viscosity_ConstantDisc(self, alpha_u, alpha_AV, beta_AV):
        self._c.update(av=4, alpha_u=alpha_u, alpha_AV=alpha_AV, beta_AV=beta_AV)

    def set_boundary_free(self):
        self._c.update(bc=0)

    def set_boundary_periodic(self):
        self._c.update(bc=1)

    def set_eos_adiabatic(self, gamma):
        self._c.update(eos=0, gamma=gamma)

    def set_eos_isothermal(self, cs):
        self._c.update(eos=1, cs0=cs)

    def set_eos_locally_isothermalLP07(self, cs0, q, r0):
        self._c.update(eos=2, cs0=cs0, eos_q=q, eos_r0=r0)

    def add_ext_force_point_mass(self, central_mass, Racc):
        self._c.update(has_point_mass=1, pm_mass=central_mass, pm_racc=Racc)

    def add_kill_sphere(self, center, radius):
        self._kill.append((tuple(center), float(radius)))

    def set_units(self, unit_system):
        self._units_G = Constants(unit_system).G()

    def set_particle_mass(self, gpart_mass):
        self._c.update(gpart_mass=gpart_mass)

    def set_tree_reduction_level(self, level):
        self._c.update(tree_reduction_level=int(level))

    def set_two_stage_search(self, enable):
        self._c.update(use_two_stage_search=int(bool(enable)))

    def set_enable_particle_reordering(self, enable):
        self._c["enable_particle_reordering"] = int(bool(enable))

    def set_particle_reordering_step_freq(self, freq):
        if int(freq) == 0:
            raise ValueError("particle_reordering_step_freq cannot be zero")
        self._c["particle_reordering_step_freq"] = int(freq)

    def set_smoothing_length_density_based(self):
        pass  # the default (and only) mode of this path

    def print_status(self):
        print("----- SPH Solver configuration (libshamb200) -----")
        for k, v in sorted(self._c.items()):
            print(f"  {k} = {v}")
        print("--------------------------------------------------")


class UnitSystem:
    """shamrock.UnitSystem (SI-based conversion factors; shamunits)"""

    def __init__(self, unit_time=1.0, unit_length=1.0, unit_mass=1.0, unit_current=1.0, unit_temperature=1.0,
                 unit_qte=1.0, unit_lumint=1.0):
        self.unit_time, self.unit_length, self.unit_mass = unit_time, unit_length, unit_mass


class Constants:
    """shamrock.Constants(unit_system): the physical constants the disc scripts read, in code units"""

    def __init__(self, unit_system):
        self.u = unit_system

    def G(self):
        # Constants<T>::Si::G = 6.6743015e-11 in the reference (shamunits/Constants.hpp:86), not CODATA's 6.67430e-11
        return 6.6743015e-11 * self.u.unit_mass * self.u.unit_time**2 / self.u.unit_length**3

    def year(self):
        return 31557600.0 / self.u.unit_time

    def au(self):
        return 149597870700.0 / self.u.unit_length

    def sol_mass(self):
        return 1.98847e30 / self.u.unit_mass


# ---- setup nodes -----------------------------------------------------------------------------------------
class _SetupNode:
    def __init__(self, pos, h):
        self.pos, self.h = pos, h


class SPHSetup:
    """model.get_setup(): modules::SPHSetup"""

    def __init__(self, model):
        self._m = model

    def make_generator_lattice_hcp(self, dr, box_min, box_max, discontinuous=True):
        # the lattice points r of the index box with box_min <= r < box_max, hpart = dr (GeneratorLatticeHCP.hpp);
        # `discontinuous` only changes the generation order, which apply_setup's Morton reordering removes
        pos = lattice.hcp_positions(dr, tuple(box_min), tuple(box_max))
        return _SetupNode(pos, np.full(len(pos), float(dr)))

    def make_combiner_add(self, parent1, parent2):
        return _SetupNode(np.concatenate([parent1.pos, parent2.pos]), np.concatenate([parent1.h, parent2.h]))

    def apply_setup(self, setup, part_reordering=True, gen_step=None, insert_step=None, msg_count_limit=None,
                    rank_comm_size_limit=None, max_msg_size=None, do_setup_log=False, use_new_setup=True,
                    speculative_balancing=False):
        if setup is None:
            raise ValueError("The setup shared pointer is empty")
        self._m._append(setup.pos, setup.h)
        # SPHSetup.cpp:202-205: modules::ParticleReordering once the particles are in place (done on the
        # device when the patch data goes up: the fields set by position in between move with their particles)
        self._m._reorder_at_push = self._m._reorder_at_push or bool(part_reordering)


# ---- analysis ----------------------------------------------------------------------------------------------
class AnalysisSodTube:
    """model.make_analysis_sodtube(...): modules::AnalysisSodTube::compute_L2_dist"""

    def __init__(self, model, sod, direction, time_val, x_ref, x_min, x_max):
        self._m, self.sod, self.direction = model, sod, np.asarray(direction, dtype=np.float64)
        self.time_val, self.x_ref, self.x_min, self.x_max = time_val, x_ref, x_min, x_max

    def compute_L2_dist(self):
        m = self._m
        if m._cfg._c["eos"] != 0:
            raise ValueError("The sod analysis is only available for adiabatic EOS")
        gamma = m._cfg._c["gamma"]
        d = m._collect()
        q = m._hfact / d["hpart"]
        rho = m._pmass * q * q * q
        P = (gamma - 1) * rho * d["uint"]
        x = d["xyz"] @ self.direction - self.x_ref
        sel = ((x + self.x_ref) > self.x_min) & ((x + self.x_ref) < self.x_max)
        if not sel.any():
            raise RuntimeError("no particle in wanted region")
        r_rho, r_vx, r_P = self.sod.get_value(self.time_val, x[sel])
        d_rho, d_P = rho[sel] - r_rho, P[sel] - r_P
        dv = d["vxyz"][sel] - r_vx[:, None] * self.direction[None, :]
        n = float(sel.sum())
        return (_math.fsum(d_rho * d_rho) / n, tuple(_math.fsum(dv[:, c] ** 2) / n for c in range(3)),
                _math.fsum(d_P * d_P) / n)


# ---- the model -----------------------------------------------------------------------------------------------
_MAIN = ("xyz", "vxyz", "axyz", "axyz_ext", "hpart", "uint", "duint", "alpha_AV", "divv", "dtdivv", "curlv",
         "soundspeed")
_NV = {"xyz": 3, "vxyz": 3, "axyz": 3, "axyz_ext": 3, "curlv": 3}


class Model:
    """shamrock.get_Model_SPH(...): shammodels::sph::Model<f64_3, Kernel>"""

    def __init__(self, context, sph_kernel, device=0, fp_mode="strict", sort_mode="bitonic"):
        if sph_kernel not in _KERNELS:
            raise ValueError(f"unknown sph kernel {sph_kernel} (this path: M4, M6)")
        self._ctx, context._model = context, self
        self._kernel, (self._kid, self._hfact, self._Rkern) = sph_kernel, _KERNELS[sph_kernel]
        self._device, self._fp_mode, self._sort_mode = device, fp_mode, sort_mode
        self._cfg = SolverConfig(sph_kernel)
        self._bmin = self._bmax = None
        self._host = {nm: np.zeros((0, 3) if nm in _NV else (0,)) for nm in _MAIN}
        self._pmass, self._cfl_cour, self._cfl_force = 0.0, 0.0, 0.0
        self._dev = self._devctx = None
        self._next_dt, self._time = None, None
        self._callbacks = []
        self._last = {}
        self._dirty, self._on_device, self._host_fresh = True, False, True
        self._reorder_at_push = False

    # -- configuration
    def gen_default_config(self):
        return SolverConfig(self._kernel)

    def set_solver_config(self, cfg):
        if self._dev is not None:
            raise RuntimeError("Cannot change solver config after scheduler is initialized")
        self._cfg = cfg

    def init_scheduler(self, crit_split, crit_merge):
        self._split, self._merge = crit_split, crit_merge

    def resize_simulation_box(self, box_min, box_max):
        self._bmin, self._bmax = tuple(float(v) for v in box_min), tuple(float(v) for v in box_max)

    def get_box_dim_fcc_3d(self, dr, xcnt, ycnt, zcnt):
        i, j, k = xcnt, ycnt, zcnt
        r = (2 * i + ((j + k) % 2), _math.sqrt(3.0) * (j + (1.0 / 3.0) * (k % 2)), 2 * _math.sqrt(6.0) * k / 3)
        return tuple(c * dr for c in r)

    def get_ideal_fcc_box(self, dr, box_min, box_max):
        return _get_ideal_hcp_box(dr, box_min, box_max)

    def set_cfl_cour(self, v):
        self._cfl_cour = float(v)
        self._reconfigure()

    def set_cfl_force(self, v):
        self._cfl_force = float(v)
        self._reconfigure()

    def set_particle_mass(self, gpart_mass):
        self._pmass = float(gpart_mass)
        self._reconfigure()

    def _reconfigure(self):
        if self._dev is not None:  # the running solver takes the new values at the next step
            c = self._dev.cfg
            c.gpart_mass, c.cfl_cour, c.cfl_force = self._pmass, self._cfl_cour, self._cfl_force
            self._dev.set_config(c)

    def get_particle_mass(self):
        return self._pmass

    def get_hfact(self):
        return self._hfact

    def rho_h(self, h):
        q = self._hfact / h
        return self._pmass * q * q * q

    def get_solver_tex(self):
        return "% solver graph introspection is not part of the B200 backend"

    def get_solver_dot_graph(self):
        return "// solver graph introspection is not part of the B200 backend"

    # -- setup (host side)
    def get_setup(self):
        return SPHSetup(self)

    def add_cube_hcp_3d(self, dr, box_min_max):
        pos = lattice.hcp_positions(dr, tuple(box_min_max[0]), tuple(box_min_max[1]))
        self._append(pos, np.full(len(pos), float(dr)))

    def _append(self, pos, h):
        self._pull()
        if self._bmin is None:
            raise RuntimeError("the box size is not set, please resize the box to the domain size")
        n_old, n_add = len(self._host["xyz"]), len(pos)
        for nm in _MAIN:
            add = np.zeros((n_add, 3) if nm in _NV else (n_add,))
            self._host[nm] = np.concatenate([self._host[nm], add])
        self._host["xyz"][n_old:] = pos
        self._host["hpart"][n_old:] = h
        self._dirty = True

    def get_total_part_count(self):
        return len(self._collect()["xyz"])

    def total_mass_to_part_mass(self, totmass):
        return totmass / self.get_total_part_count()

    def _in_box(self, box_min, box_max):
        x = self._host["xyz"]
        sel = np.ones(len(x), dtype=bool)
        for c in range(3):
            sel &= (box_min[c] <= x[:, c]) & (x[:, c] < box_max[c])
        return sel

    def set_value_in_a_box(self, field_name, field_type, val, box_min, box_max, ivar=0):
        self._pull()
        f = self._host[field_name]
        sel = self._in_box(box_min, box_max)
        if field_name in _NV:
            f[sel] = np.asarray(val, dtype=np.float64)
        else:
            f[sel] = val
        self._dirty = True

    def set_value_in_sphere(self, field_name, field_type, val, center, radius):
        self._pull()
        d = self._host["xyz"] - np.asarray(center, dtype=np.float64)
        self._host[field_name][np.einsum("ij,ij->i", d, d) < radius * radius] = val
        self._dirty = True

    def add_kernel_value(self, field_name, field_type, val, center, h_ker):
        """f += val * W_3d(|r - center|, h_ker) with the model's kernel (Model.hpp:740-768)"""
        self._pull()
        r = np.linalg.norm(self._host["xyz"] - np.asarray(center, dtype=np.float64), axis=1)
        q = r / h_ker
        if self._kernel == "M4":
            f = np.where(q < 1, 0.25 * (2 - q) ** 3 - (1 - q) ** 3, np.where(q < 2, 0.25 * (2 - q) ** 3, 0.0))
            norm = 1 / _math.pi
        else:
            t1, t2, t3 = (3 - q) ** 5, -6 * (2 - q) ** 5, 15 * (1 - q) ** 5
            f = np.where(q < 1, t1 + t2 + t3, np.where(q < 2, t1 + t2, np.where(q < 3, t1, 0.0)))
            norm = 1 / (120 * _math.pi)
        self._host[field_name] += val * norm * f / (h_ker * h_ker * h_ker)
        self._dirty = True

    def get_sum(self, field_name, field_type):
        f = self._collect()[field_name]
        return f.sum(axis=0) if f.ndim == 2 else float(f.sum())

    def get_closest_part_to(self, pos):
        x = self._collect()["xyz"]
        d = x - np.asarray(pos, dtype=np.float64)
        return tuple(x[int(np.argmin(np.einsum("ij,ij->i", d, d)))])

    # -- device residency
    def _make_device_model(self):
        c = _capi.default_config()
        kv = dict(self._cfg._c)
        kv.update(kernel=self._kid, gpart_mass=self._pmass if self._pmass else kv.get("gpart_mass", 0.0),
                  cfl_cour=self._cfl_cour, cfl_force=self._cfl_force, constant_G=self._cfg._units_G)
        for k, v in kv.items():
            cur = getattr(c, k)
            setattr(c, k, int(v) if isinstance(cur, int) else float(v))
        c.sort_mode = _capi.SORT_MODES[self._sort_mode]
        c.fp_mode = _capi.FP_MODES[self._fp_mode]
        c.keep_step_data = 0
        for i, (ctr, r) in enumerate(self._cfg._kill):
            for d in range(3):
                c.kill_center[i][d] = ctr[d]
            c.kill_radius[i] = r
        c.n_kill_spheres = len(self._cfg._kill)
        self._devctx = _capi.Context(self._device)
        self._dev = _capi.Model(self._devctx, c)
        self._dev.set_box(self._bmin, self._bmax, (1, 1, 1))

    def _push(self):
        """host patch data -> device (first timestep, or after a host-side edit)"""
        if self._dev is None:
            self._make_device_model()
            self._dev.push_particles(self._host["xyz"], self._host["vxyz"], self._host["hpart"], self._host["uint"])
            first = True
        else:
            first = False
        if not first or any(np.any(self._host[nm]) for nm in ("axyz", "duint", "alpha_AV")):
            if self._dev.patch_size(0) != len(self._host["xyz"]):
                raise RuntimeError("particles cannot be added once the simulation has started")
            for nm in _MAIN:
                self._dev.set_field(0, nm, self._host[nm])
        if first and self._reorder_at_push:
            self._dev.reorder_particles()
            self._host_fresh = False  # the host copy is in generation order
        self._dirty, self._on_device = False, True

    def _pull(self):
        """device -> host before a host-side read or edit"""
        if self._dev is not None and getattr(self, "_on_device", False) and not getattr(self, "_host_fresh", False):
            for nm in _MAIN:
                self._host[nm] = self._dev.get(0, nm)
            self._host_fresh = True

    def _collect(self):
        self._pull()
        return self._host

    # -- time stepping
    def set_next_dt(self, dt):
        self._next_dt = float(dt)

    def get_time(self):
        return self._dev.state()["time"] if self._dev else 0.0

    def get_dt(self):
        return self._dev.state()["dt"] if self._dev else 0.0

    def add_timestep_callback(self, step_begin=None, step_end=None):
        self._callbacks.append((step_begin, step_end))

    def evolve_once(self):
        if getattr(self, "_dirty", True) or self._dev is None:
            self._push()
        if self._next_dt is not None:
            self._dev.set_next_dt(self._next_dt)
            self._next_dt = None
        for b, _ in self._callbacks:
            if b:
                b()
        self._last = self._dev.evolve_once()
        self._host_fresh = False
        for _, e in self._callbacks:
            if e:
                e()
        return self._last

    def timestep(self):
        return self.evolve_once()

    def evolve_until(self, target_time, niter_max=-1):
        """Solver::evolve_until (Solver.hpp:305-370): dt clipped to land on the target time"""
        if getattr(self, "_dirty", True) or self._dev is None:
            self._push()
        n = 0
        while self.get_time() < target_time:
            st = self._dev.state()
            if st["time"] > target_time:
                raise ValueError("the target time is higher than the current time")
            if self._next_dt is None and st["time"] + st["dt"] > target_time:
                self.set_next_dt(target_time - st["time"])
            elif self._next_dt is not None and st["time"] + self._next_dt > target_time:
                self.set_next_dt(target_time - st["time"])
            self.evolve_once()
            n += 1
            if 0 <= niter_max <= n:
                break
        return n

    # -- Phantom dumps / VTK (pySPHModel.cpp:1352-1379)
    def do_vtk_dump(self, filename, add_patch_world_id):
        if getattr(self, "_dirty", True) or self._dev is None:
            self._push()
        self._dev.vtk_dump(filename, add_patch_world_id)

    def make_phantom_dump(self):
        """a PhantomDump of the current state (held in a temporary file until save_dump copies it)"""
        import tempfile

        if getattr(self, "_dirty", True) or self._dev is None:
            self._push()
        tmp = tempfile.NamedTemporaryFile(prefix="shamb200_", suffix=".phdump", delete=False)
        tmp.close()
        self._dev.phantom_dump(tmp.name)
        return PhantomDump(tmp.name, owned=True)

    def gen_config_from_phantom_dump(self, dump, bypass_error=False):
        c = _capi.phantom_gen_config(dump._fname, bypass_error=bypass_error)
        cfg = SolverConfig(self._kernel)
        cfg._c.update(eos=c.eos, gamma=c.gamma, cs0=c.cs0, eos_q=c.eos_q, eos_r0=c.eos_r0, av=c.av,
                      alpha_min=c.alpha_min, alpha_max=c.alpha_max, sigma_decay=c.sigma_decay, alpha_u=c.alpha_u,
                      beta_AV=c.beta_AV, bc=c.bc, gpart_mass=c.gpart_mass)
        cfg._cfl = (c.cfl_cour, c.cfl_force)
        return cfg

    def init_from_phantom_dump(self, dump, hpart_fact_load=1.0):
        """particles, box and time of the dump; call after set_solver_config + init_scheduler like the reference"""
        if self._dev is not None:
            raise RuntimeError("init_from_phantom_dump needs a model that has not started")
        if hasattr(self._cfg, "_cfl") and not self._cfl_cour:
            self._cfl_cour, self._cfl_force = self._cfg._cfl
        if not self._pmass:
            self._pmass = self._cfg._c.get("gpart_mass", 0.0)
        self._bmin, self._bmax = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)  # replaced by the dump's box
        self._make_device_model()
        self._dev.init_from_phantom_dump(dump._fname, hpart_fact_load)
        info = self._dev.patch_info(0)
        self._bmin, self._bmax = info["lo"], info["hi"]
        self._dirty, self._on_device, self._host_fresh = False, True, False
        self._pull()

    # -- checkpoint / restart (Model::dump / Model::load_from_dump, Model.hpp:906-995)
    def dump(self, fname):
        if getattr(self, "_dirty", True) or self._dev is None:
            self._push()
        self._dev.dump(fname)

    def load_from_dump(self, fname):
        """configuration, patches and state come from the dump; one patch list entry per patch of the dump"""
        if self._dev is None:
            if self._bmin is None:
                self._bmin, self._bmax = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
            self._make_device_model()
        self._dev.load_dump(fname)
        if self._dev.patch_count != 1:
            raise RuntimeError("this Python surface drives one patch; load multi-patch dumps through _capi.Model")
        self._dirty, self._on_device, self._host_fresh = False, True, False
        self._pull()

    def get_total_part_count_device(self):
        return self._dev.total_part_count() if self._dev else len(self._host["xyz"])

    def solver_logs_last_rate(self):
        return self._last.get("rate", 0.0)

    def solver_logs_last_obj_count(self):
        return int(self._last.get("npart", 0))

    def solver_logs_last_system_metrics(self):
        return {}

    def make_analysis_sodtube(self, sod, direction, time_val, x_ref, x_min, x_max):
        return AnalysisSodTube(self, sod, direction, time_val, x_ref, x_min, x_max)


class PhantomDump:
    """shamrock.PhantomDump (pyPhantomDump.cpp:29-46): a dump file on disk, read through libshamb200"""

    def __init__(self, fname, owned=False):
        self._fname, self._owned = str(fname), owned

    def __del__(self):
        if getattr(self, "_owned", False):
            import os

            try:
                os.unlink(self._fname)
            except OSError:
                pass

    def save_dump(self, fname):
        _capi.phantom_copy(self._fname, fname)  # from_file + gen_file + write_to_file

    def read_header_float(self, s):
        v = _capi.phantom_header_float(self._fname, s)
        if v is None:
            raise RuntimeError("the entry cannot be found : " + s)
        return v

    def read_header_int(self, s):
        v = _capi.phantom_header_int(self._fname, s)
        if v is None:
            raise RuntimeError("the entry cannot be found")
        return v

    def print_state(self):
        print("--- dump state ---")
        print("file =", self._fname, " nparttot =", _capi.phantom_header_int(self._fname, "nparttot"))
        print("------------------")


def load_phantom_dump(fname):
    d = PhantomDump(fname)
    d.read_header_int("nparttot")  # parses the file: a damaged dump fails here, as in the reference
    return d


def compare_phantom_dumps(dump_1, dump_2):
    return _capi.phantom_compare(dump_1._fname, dump_2._fname) == 0


def get_Model_SPH(context, vector_type="f64_3", sph_kernel="M4", device=0, fp_mode="strict", sort_mode="bitonic"):
    """shamrock.get_Model_SPH(context=ctx, vector_type="f64_3", sph_kernel="M4" | "M6").  Backend options
    (not in the reference): device ordinal; fp_mode "strict" (bit-identical to the reference's expression
    order) or "fast"; sort_mode "bitonic" (the reference's network incl. its order of equal Morton codes) or
    "radix" (faster, equal codes keep the input order)."""
    if vector_type != "f64_3":
        raise ValueError("unknown combination of representation and kernel (this path: f64_3 with M4 or M6)")
    return Model(context, sph_kernel, device=device, fp_mode=fp_mode, sort_mode=sort_mode)
