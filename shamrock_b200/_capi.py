"""ctypes binding of libshamb200.so (include/shamb200.h).  No CPU fallback: importing works without a
GPU (symbols can be inspected), every compute entry point needs a CUDA device."""
import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libshamb200.so")

KERNELS = {"M4": 0, "M6": 1}
SORT_MODES = {"bitonic": 0, "radix": 1}
EOS = {"adiabatic": 0, "isothermal": 1, "locally_isothermal_lp07": 2}
AV = {"none": 0, "constant": 1, "varying_mm97": 2, "varying_cd10": 3, "constant_disc": 4}
BC = {"free": 0, "periodic": 1}
FP_MODES = {"strict": 0, "fast": 1}


class ShamB200Error(RuntimeError):
    pass


class TreeView(C.Structure):
    _fields_ = [
        ("obj_cnt", C.c_uint32), ("morton_count", C.c_uint32), ("leaf_count", C.c_uint32),
        ("int_count", C.c_uint32), ("bmin", C.c_double * 3), ("bmax", C.c_double * 3),
        ("d_sorted_morton", C.c_void_p), ("d_sort_index_map", C.c_void_p),
        ("d_reduc_index_map", C.c_void_p), ("d_reduced_morton", C.c_void_p),
        ("d_lchild_id", C.c_void_p), ("d_rchild_id", C.c_void_p), ("d_endrange", C.c_void_p),
        ("d_lchild_flag", C.c_void_p), ("d_rchild_flag", C.c_void_p),
        ("d_aabb_min", C.c_void_p), ("d_aabb_max", C.c_void_p),
    ]


class CsrView(C.Structure):
    _fields_ = [
        ("obj_cnt", C.c_uint32), ("sum_neigh_cnt", C.c_uint32), ("d_cnt_neigh", C.c_void_p),
        ("d_scanned_cnt", C.c_void_p), ("d_index_neigh_map", C.c_void_p),
    ]


class MergedFields(C.Structure):
    """shamb200_merged_fields: device pointers of the merged (real + ghost) fields of one patch"""
    _fields_ = [
        ("obj_cnt", C.c_uint32), ("real_cnt", C.c_uint32), ("d_xyz", C.c_void_p), ("stride_dbl", C.c_size_t),
        ("d_hpart", C.c_void_p), ("d_vxyz", C.c_void_p), ("d_uint", C.c_void_p), ("d_axyz", C.c_void_p),
        ("d_omega", C.c_void_p), ("d_pressure", C.c_void_p), ("d_soundspeed", C.c_void_p),
        ("d_alpha_AV", C.c_void_p),
    ]


class SolverConfig(C.Structure):
    _fields_ = [
        ("kernel", C.c_int32), ("eos", C.c_int32), ("av", C.c_int32), ("bc", C.c_int32),
        ("gpart_mass", C.c_double),
        ("gamma", C.c_double), ("cs0", C.c_double), ("eos_q", C.c_double), ("eos_r0", C.c_double),
        ("alpha_u", C.c_double), ("alpha_AV", C.c_double), ("beta_AV", C.c_double),
        ("alpha_min", C.c_double), ("alpha_max", C.c_double), ("sigma_decay", C.c_double),
        ("cfl_cour", C.c_double), ("cfl_force", C.c_double), ("cfl_multiplier_stiffness", C.c_double),
        ("htol_up_coarse_cycle", C.c_double), ("htol_up_fine_cycle", C.c_double), ("epsilon_h", C.c_double),
        ("h_iter_per_subcycles", C.c_uint32), ("h_max_subcycles_count", C.c_uint32),
        ("tree_reduction_level", C.c_uint32),
        ("use_two_stage_search", C.c_int32), ("combined_dtdiv_divcurlv_compute", C.c_int32),
        ("sort_mode", C.c_int32), ("has_point_mass", C.c_int32),
        ("pm_mass", C.c_double), ("pm_racc", C.c_double), ("constant_G", C.c_double),
        ("n_kill_spheres", C.c_int32), ("keep_step_data", C.c_int32),
        ("fp_mode", C.c_int32), ("enable_particle_reordering", C.c_int32),
        ("kill_center", (C.c_double * 3) * 4), ("kill_radius", C.c_double * 4),
        ("particle_reordering_step_freq", C.c_uint64),
    ]


class Iface(C.Structure):
    _fields_ = [("sender", C.c_uint32), ("receiver", C.c_uint32), ("ioff", C.c_int32 * 3),
                ("offset", C.c_double * 3), ("cut_lo", C.c_double * 3), ("cut_hi", C.c_double * 3)]


# every symbol include/shamb200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "shamb200_last_error", "shamb200_build_info", "shamb200_launch_count", "shamb200_reset_launch_count",
    "shamb200_ctx_create", "shamb200_ctx_destroy", "shamb200_ctx_stream", "shamb200_ctx_synchronize",
    "shamb200_tree_build", "shamb200_tree_build_auto_bbox", "shamb200_tree_field_max",
    "shamb200_neigh_cache_build", "shamb200_neigh_cache_stats", "shamb200_h_iterate", "shamb200_h_iterate_loop", "shamb200_compute_omega",
    "shamb200_solver_config_default", "shamb200_model_create", "shamb200_model_destroy",
    "shamb200_model_set_config", "shamb200_model_set_box", "shamb200_nccl_unique_id",
    "shamb200_model_init_comm", "shamb200_model_push_particles", "shamb200_model_patch_count",
    "shamb200_model_patch_is_local", "shamb200_model_patch_size", "shamb200_model_get",
    "shamb200_model_set_field", "shamb200_model_reorder_particles", "shamb200_model_evolve_once",
    "shamb200_model_evolve_once_host",
    "shamb200_host_register", "shamb200_host_unregister", "shamb200_model_host_traffic",
    "shamb200_model_host_step_info", "shamb200_model_list_tolerance",
    "shamb200_model_search_stats", "shamb200_model_state", "shamb200_model_conservation",
    "shamb200_model_add_lattice_hcp", "shamb200_model_add_disc_lattice", "shamb200_model_add_disc_mc", "shamb200_model_set_value_in_a_box",
    "shamb200_model_set_value_in_sphere", "shamb200_model_add_kernel_value", "shamb200_model_get_sum",
    "shamb200_model_total_part_count", "shamb200_model_set_particle_mass",
    "shamb200_model_dump", "shamb200_model_load_dump",
    "shamb200_model_phantom_dump", "shamb200_model_init_from_phantom_dump", "shamb200_model_vtk_dump",
    "shamb200_phantom_gen_config", "shamb200_phantom_copy", "shamb200_phantom_header_float",
    "shamb200_phantom_header_int", "shamb200_phantom_compare", "shamb200_model_init_scheduler", "shamb200_model_scheduler_step", "shamb200_model_split_patch",
    "shamb200_model_merge_patches", "shamb200_model_migrate_patch", "shamb200_model_patch_info",
    "shamb200_model_scheduler_log",
    "shamb200_model_set_next_dt", "shamb200_model_set_time", "shamb200_model_set_cfl_multiplier",
    "shamb200_model_stage_times", "shamb200_plan_patch_grid", "shamb200_plan_interfaces",
    "shamb200_microbench", "shamb200_hilbert_index", "shamb200_plan_load_balance",
    "shamb200_model_set_patch_owners", "shamb200_model_patch_coords",
    "shamb200_compute_eos", "shamb200_update_divv_curlv", "shamb200_update_dtdivv", "shamb200_update_viscosity",
    "shamb200_update_derivs", "shamb200_vsig_cfl", "shamb200_leapfrog_predict", "shamb200_leapfrog_correct",
]

HOST_FIELDS = (("xyz", 3), ("vxyz", 3), ("axyz", 3), ("axyz_ext", 3), ("hpart", 1), ("uint", 1), ("duint", 1),
               ("alpha_AV", 1), ("divv", 1), ("dtdivv", 1), ("curlv", 3), ("soundspeed", 1))


class HostPatchData(C.Structure):
    """shamb200_host_patchdata: host pointers of the main-layout fields of one patch"""
    _fields_ = [("n", C.c_uint64)] + [(nm if nm != "uint" else "uint_", C.c_void_p) for nm, _ in HOST_FIELDS]


_lib = None


def lib():
    """Load libshamb200.so; fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ShamB200Error(
                f"{LIB_PATH} is missing: build it with `python -m shamrock_b200.build` "
                "(nvcc, sm_100a).  shamrock_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        L.shamb200_last_error.restype = C.c_char_p
        L.shamb200_build_info.restype = C.c_char_p
        L.shamb200_launch_count.restype = C.c_uint64
        L.shamb200_ctx_stream.restype = C.c_void_p
        L.shamb200_model_get.restype = C.c_int64
        L.shamb200_model_patch_count.restype = C.c_uint32
        L.shamb200_model_patch_size.restype = C.c_uint32
        L.shamb200_hilbert_index.restype = C.c_uint64
        L.shamb200_hilbert_index.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.shamb200_ctx_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.shamb200_model_create.argtypes = [C.c_void_p, C.POINTER(SolverConfig), C.POINTER(C.c_void_p)]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise ShamB200Error(f"[{rc}] " + lib().shamb200_last_error().decode())


def hilbert_index(x, y, z):
    """Hilbert index of a cell of the 2^21-per-axis patch grid (host only)"""
    return int(lib().shamb200_hilbert_index(int(x), int(y), int(z)))


def plan_load_balance(coord_min, load, world_size):
    """patch -> rank along the Hilbert curve (HilbertLoadBalance + load_balance of the reference, host only).
    Returns (owner[npatch], strategy) with strategy 'psweep' or 'round robin'."""
    c = np.ascontiguousarray(coord_min, dtype=np.uint64).reshape(-1, 3)
    l = np.ascontiguousarray(load, dtype=np.uint64)
    assert len(c) == len(l)
    owner = np.zeros(len(l), dtype=np.int32)
    strat = C.c_int(0)
    check(lib().shamb200_plan_load_balance(
        C.c_uint32(len(l)), c.ctypes.data_as(C.c_void_p), l.ctypes.data_as(C.c_void_p), int(world_size),
        owner.ctypes.data_as(C.c_void_p), C.byref(strat)))
    return owner, ("psweep", "round robin")[strat.value]


def default_config():
    cfg = SolverConfig()
    lib().shamb200_solver_config_default(C.byref(cfg))
    return cfg


def launch_count():
    return int(lib().shamb200_launch_count())


def reset_launch_count():
    lib().shamb200_reset_launch_count()


class Context:
    """One per GPU (shamb200_ctx)."""

    def __init__(self, device=0, stream=None):
        h = C.c_void_p()
        check(lib().shamb200_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self.h = h
        self.device = device
        self._models = weakref.WeakSet()  # models running on this context: destroyed before it

    def close(self):
        if getattr(self, "h", None):
            # the garbage collector finalises a model and its context in any order: a model must never
            # outlive the stream it runs on
            for m in list(getattr(self, "_models", ())):
                m.close()
            lib().shamb200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def microbench(self, what):
        """'fp64' -> TFLOP/s of FP64 FMA chains, 'copy' -> GB/s of a streaming copy (read + write)"""
        out = C.c_double()
        code = {"fp64": 0, "copy": 1}.get(what, what)  # 10 + p / 20 + p: record gathers (microbench.cu)
        check(lib().shamb200_microbench(self.h, int(code), C.byref(out)))
        return out.value

    @property
    def stream(self):
        return lib().shamb200_ctx_stream(self.h)

    def synchronize(self):
        check(lib().shamb200_ctx_synchronize(self.h))

    # ---- stage-level entry points on device pointers (torch tensors carry the memory) ----------
    def tree_build(self, xyz_t, obj_cnt, bmin=None, bmax=None, reduction_level=3, sort_mode="bitonic",
                   stride_dbl=3):
        tv = TreeView()
        if bmin is None:
            check(lib().shamb200_tree_build_auto_bbox(
                self.h, C.c_void_p(xyz_t.data_ptr()), C.c_size_t(stride_dbl), C.c_uint32(obj_cnt),
                C.c_uint32(reduction_level), SORT_MODES[sort_mode], C.byref(tv)))
        else:
            b0 = (C.c_double * 3)(*bmin)
            b1 = (C.c_double * 3)(*bmax)
            check(lib().shamb200_tree_build(
                self.h, C.c_void_p(xyz_t.data_ptr()), C.c_size_t(stride_dbl), C.c_uint32(obj_cnt), b0, b1,
                C.c_uint32(reduction_level), SORT_MODES[sort_mode], C.byref(tv)))
        return tv

    def tree_field_max(self, tv, field_t, scale, out_t):
        check(lib().shamb200_tree_field_max(self.h, C.byref(tv), C.c_void_p(field_t.data_ptr()),
                                            C.c_double(scale), C.c_void_p(out_t.data_ptr())))

    def neigh_cache_build(self, tv, xyz_t, h_t, rint_t, obj_cnt, Rkern, htol, two_stage=True, stride_dbl=3):
        cv = CsrView()
        check(lib().shamb200_neigh_cache_build(
            self.h, C.byref(tv), C.c_void_p(xyz_t.data_ptr()), C.c_size_t(stride_dbl),
            C.c_void_p(h_t.data_ptr()), C.c_void_p(rint_t.data_ptr()), C.c_uint32(obj_cnt),
            C.c_double(Rkern), C.c_double(htol), int(two_stage), C.byref(cv)))
        return cv

    def neigh_cache_stats(self):
        out = (C.c_uint64 * 6)()
        check(lib().shamb200_neigh_cache_stats(self.h, out))
        return dict(K=out[0], pair_tests=out[1], attempts=out[2], over_groups=out[3], frontier_cap=out[4],
                    candidate_cap=out[5])

    def h_iterate(self, kernel, cv, xyz_t, h_old_t, h_new_t, eps_t, pmass, h_evol_max, h_evol_iter_max,
                  stride_dbl=3):
        check(lib().shamb200_h_iterate(
            self.h, KERNELS[kernel], C.byref(cv), C.c_void_p(xyz_t.data_ptr()), C.c_size_t(stride_dbl),
            C.c_void_p(h_old_t.data_ptr()), C.c_void_p(h_new_t.data_ptr()), C.c_void_p(eps_t.data_ptr()),
            C.c_double(pmass), C.c_double(h_evol_max), C.c_double(h_evol_iter_max)))

    def h_iterate_loop(self, kernel, cv, xyz_t, h_old_t, h_new_t, eps_t, pmass, h_evol_max, h_evol_iter_max,
                       epsilon_h, max_sweeps, stride_dbl=3):
        out = (C.c_double * 3)()
        check(lib().shamb200_h_iterate_loop(
            self.h, KERNELS[kernel], C.byref(cv), C.c_void_p(xyz_t.data_ptr()), C.c_size_t(stride_dbl),
            C.c_void_p(h_old_t.data_ptr()), C.c_void_p(h_new_t.data_ptr()), C.c_void_p(eps_t.data_ptr()),
            C.c_double(pmass), C.c_double(h_evol_max), C.c_double(h_evol_iter_max), C.c_double(epsilon_h),
            C.c_uint32(max_sweeps), out))
        return dict(max_eps=out[0], min_eps=out[1], sweeps=int(out[2]))

    def compute_omega(self, kernel, cv, xyz_t, h_t, omega_t, pmass, stride_dbl=3):
        check(lib().shamb200_compute_omega(
            self.h, KERNELS[kernel], C.byref(cv), C.c_void_p(xyz_t.data_ptr()), C.c_size_t(stride_dbl),
            C.c_void_p(h_t.data_ptr()), C.c_void_p(omega_t.data_ptr()), C.c_double(pmass)))


    # ---- SPH modules on merged patch data (torch tensors on this context's device) ------------------
    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    @classmethod
    def merged_fields(cls, obj_cnt, real_cnt, xyz, hpart, vxyz=None, uint=None, axyz=None, omega=None,
                      pressure=None, soundspeed=None, alpha_AV=None, stride_dbl=3):
        """A shamb200_merged_fields view of device tensors (kept alive by the caller)."""
        f = MergedFields()
        f.obj_cnt, f.real_cnt, f.stride_dbl = int(obj_cnt), int(real_cnt), int(stride_dbl)
        for name, t in (("d_xyz", xyz), ("d_hpart", hpart), ("d_vxyz", vxyz), ("d_uint", uint), ("d_axyz", axyz),
                        ("d_omega", omega), ("d_pressure", pressure), ("d_soundspeed", soundspeed),
                        ("d_alpha_AV", alpha_AV)):
            setattr(f, name, t.data_ptr() if t is not None else None)
        return f

    def compute_eos(self, kernel, eos, f, pmass, gamma, cs0, eos_q, eos_r0, pressure_t, soundspeed_t):
        check(lib().shamb200_compute_eos(
            self.h, KERNELS[kernel], EOS[eos], C.byref(f), C.c_double(pmass), C.c_double(gamma), C.c_double(cs0),
            C.c_double(eos_q), C.c_double(eos_r0), self._p(pressure_t), self._p(soundspeed_t)))

    def update_divv_curlv(self, kernel, cv, f, pmass, divv_t, curlv_t=None):
        check(lib().shamb200_update_divv_curlv(
            self.h, KERNELS[kernel], C.byref(cv), C.byref(f), C.c_double(pmass), self._p(divv_t), self._p(curlv_t)))

    def update_dtdivv(self, kernel, cv, f, pmass, dtdivv_t, also_divv_curlv=False, divv_t=None, curlv_t=None):
        check(lib().shamb200_update_dtdivv(
            self.h, KERNELS[kernel], C.byref(cv), C.byref(f), C.c_double(pmass), int(also_divv_curlv),
            self._p(divv_t), self._p(curlv_t), self._p(dtdivv_t)))

    def update_viscosity(self, av, n, dt, sigma_decay, alpha_min, alpha_max, divv_t, curlv_t, dtdivv_t, cs_t, h_t,
                         alpha_t, alpha_out_t):
        check(lib().shamb200_update_viscosity(
            self.h, AV[av], C.c_uint32(n), C.c_double(dt), C.c_double(sigma_decay), C.c_double(alpha_min),
            C.c_double(alpha_max), self._p(divv_t), self._p(curlv_t), self._p(dtdivv_t), self._p(cs_t), self._p(h_t),
            self._p(alpha_t), self._p(alpha_out_t)))

    def update_derivs(self, kernel, av, cv, f, pmass, alpha_u, alpha_AV, beta_AV, axyz_ext_t, axyz_t, duint_t):
        check(lib().shamb200_update_derivs(
            self.h, KERNELS[kernel], AV[av], C.byref(cv), C.byref(f), C.c_double(pmass), C.c_double(alpha_u),
            C.c_double(alpha_AV), C.c_double(beta_AV), self._p(axyz_ext_t), self._p(axyz_t), self._p(duint_t)))

    def vsig_cfl(self, kernel, cv, f, axyz_t, C_cour, C_force, vsig_t, cfl_dt_t):
        out = C.c_double()
        check(lib().shamb200_vsig_cfl(
            self.h, KERNELS[kernel], C.byref(cv), C.byref(f), self._p(axyz_t), C.c_double(C_cour), C.c_double(C_force),
            self._p(vsig_t), self._p(cfl_dt_t), C.byref(out)))
        return out.value

    def leapfrog_predict(self, n, dt, xyz_t, vxyz_t, axyz_t, uint_t, duint_t):
        check(lib().shamb200_leapfrog_predict(
            self.h, C.c_uint32(n), C.c_double(dt), self._p(xyz_t), self._p(vxyz_t), self._p(axyz_t), self._p(uint_t),
            self._p(duint_t)))

    def leapfrog_correct(self, n, half_dt, vxyz_t, axyz_t, axyz_old_t, uint_t, duint_t, duint_old_t):
        out = (C.c_double * 2)()
        check(lib().shamb200_leapfrog_correct(
            self.h, C.c_uint32(n), C.c_double(half_dt), self._p(vxyz_t), self._p(axyz_t), self._p(axyz_old_t),
            self._p(uint_t), self._p(duint_t), self._p(duint_old_t), out))
        return out[0], out[1]


_U32 = {"sorted_morton", "sort_index_map", "reduc_index_map", "reduced_morton", "lchild_id", "rchild_id",
        "endrange", "cnt_neigh", "scanned_cnt", "index_neigh_map"}
_U8 = {"lchild_flag", "rchild_flag"}
_VEC3 = {"xyz", "vxyz", "axyz", "axyz_ext", "curlv", "mxyz", "g_v", "g_a", "aabb_min", "aabb_max"}


class Model:
    """shamb200_model: device-resident patches + evolve_once (host buffers in / out)."""

    def __init__(self, ctx, cfg):
        self.ctx = ctx
        self.cfg = cfg
        h = C.c_void_p()
        check(lib().shamb200_model_create(ctx.h, C.byref(cfg), C.byref(h)))
        self.h = h
        ctx._models.add(self)

    def close(self):
        if getattr(self, "h", None):
            lib().shamb200_model_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_config(self, cfg):
        self.cfg = cfg
        check(lib().shamb200_model_set_config(self.h, C.byref(cfg)))

    def set_box(self, bmin, bmax, grid=(1, 1, 1)):
        check(lib().shamb200_model_set_box(self.h, (C.c_double * 3)(*bmin), (C.c_double * 3)(*bmax),
                                           *[C.c_uint32(g) for g in grid]))

    def patch_coords(self):
        """coord_min of every patch on the 2^21 integer grid, [npatch, 3]"""
        n = self.patch_count
        out = np.zeros((n, 3), dtype=np.uint64)
        check(lib().shamb200_model_patch_coords(self.h, C.c_uint32(n), out.ctypes.data_as(C.c_void_p)))
        return out

    def set_patch_owners(self, owner):
        """patch -> rank table (before any particle is pushed), e.g. from plan_load_balance"""
        o = np.ascontiguousarray(owner, dtype=np.int32)
        check(lib().shamb200_model_set_patch_owners(self.h, C.c_uint32(len(o)), o.ctypes.data_as(C.c_void_p)))

    def init_comm(self, rank, world, nccl_id):
        buf = (C.c_char * 128).from_buffer_copy(bytes(nccl_id))
        check(lib().shamb200_model_init_comm(self.h, int(rank), int(world), buf))

    def push_particles(self, xyz, vxyz, h, u):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        n = len(xyz)

        def p(a, nv):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64).reshape(n * nv)
            keep.append(a)
            return a.ctypes.data_as(C.c_void_p)

        keep = [xyz]
        check(lib().shamb200_model_push_particles(self.h, C.c_uint64(n), xyz.ctypes.data_as(C.c_void_p),
                                                  p(vxyz, 3), p(h, 1), p(u, 1)))

    # -- checkpoint / restart (dump.cu)
    def dump(self, fname):
        check(lib().shamb200_model_dump(self.h, str(fname).encode()))

    def load_dump(self, fname):
        check(lib().shamb200_model_load_dump(self.h, str(fname).encode()))

    # -- Phantom dumps / legacy VTK (io_formats.cu)
    def phantom_dump(self, fname):
        check(lib().shamb200_model_phantom_dump(self.h, str(fname).encode()))

    def init_from_phantom_dump(self, fname, hpart_fact_load=1.0):
        kept = C.c_uint64(0)
        check(lib().shamb200_model_init_from_phantom_dump(self.h, str(fname).encode(), C.c_double(hpart_fact_load),
                                                          C.byref(kept)))
        return kept.value

    def vtk_dump(self, fname, add_patch_world_id=True):
        check(lib().shamb200_model_vtk_dump(self.h, str(fname).encode(), C.c_int(1 if add_patch_world_id else 0)))

    # -- patch scheduler (scheduler.cu)
    def init_scheduler(self, crit_split, crit_merge, step_freq=0):
        check(lib().shamb200_model_init_scheduler(self.h, C.c_uint64(int(crit_split)), C.c_uint64(int(crit_merge)),
                                                  C.c_uint32(step_freq)))

    def scheduler_step(self, do_split_merge=True, do_load_balancing=True):
        check(lib().shamb200_model_scheduler_step(self.h, int(do_split_merge), int(do_load_balancing)))
        return self.scheduler_log()

    def split_patch(self, ip):
        check(lib().shamb200_model_split_patch(self.h, C.c_uint32(ip)))

    def merge_patches(self, ip0):
        check(lib().shamb200_model_merge_patches(self.h, C.c_uint32(ip0)))

    def migrate_patch(self, ip, new_owner):
        check(lib().shamb200_model_migrate_patch(self.h, C.c_uint32(ip), int(new_owner)))

    def patch_info(self, ip):
        o, b = (C.c_uint64 * 8)(), (C.c_double * 6)()
        check(lib().shamb200_model_patch_info(self.h, C.c_uint32(ip), o, b))
        return dict(id=int(o[0]), coord_min=tuple(o[1:4]), coord_max=tuple(o[4:7]), owner=int(o[7]),
                    lo=tuple(b[0:3]), hi=tuple(b[3:6]))

    def scheduler_log(self):
        o = (C.c_double * 8)()
        check(lib().shamb200_model_scheduler_log(self.h, o))
        return dict(splits=int(o[0]), merges=int(o[1]), moves=int(o[2]), moved_objects=int(o[3]), npatch=int(o[4]),
                    max_rank_load=int(o[5]), mean_rank_load=o[6], imbalance=o[7])

    # -- initial conditions generated on the device (setup.cu)
    def add_lattice_hcp(self, dr, box_min, box_max):
        n = C.c_uint64()
        check(lib().shamb200_model_add_lattice_hcp(self.h, C.c_double(dr), (C.c_double * 3)(*box_min),
                                                   (C.c_double * 3)(*box_max), C.byref(n)))
        return int(n.value)

    def add_disc_lattice(self, dr, r_in, r_out, zcut):
        n = C.c_uint64()
        check(lib().shamb200_model_add_disc_lattice(self.h, C.c_double(dr), C.c_double(r_in), C.c_double(r_out),
                                                    C.c_double(zcut), C.byref(n)))
        return int(n.value)

    def add_disc_mc(self, npart, seed, r_in, r_out, p, q, H_r_in, disc_mass):
        n = C.c_uint64()
        check(lib().shamb200_model_add_disc_mc(self.h, C.c_uint64(npart), C.c_uint64(seed), C.c_double(r_in),
                                               C.c_double(r_out), C.c_double(p), C.c_double(q), C.c_double(H_r_in),
                                               C.c_double(disc_mass), C.byref(n)))
        return int(n.value)

    def set_value_in_a_box(self, field, val, box_min, box_max, ivar=0):
        vals = np.atleast_1d(np.asarray(val, dtype=np.float64))
        for k, v in enumerate(vals):
            check(lib().shamb200_model_set_value_in_a_box(self.h, field.encode(), int(ivar + k if len(vals) > 1 else ivar),
                                                          C.c_double(v), (C.c_double * 3)(*box_min),
                                                          (C.c_double * 3)(*box_max)))

    def set_value_in_sphere(self, field, val, center, radius):
        check(lib().shamb200_model_set_value_in_sphere(self.h, field.encode(), C.c_double(val),
                                                       (C.c_double * 3)(*center), C.c_double(radius)))

    def add_kernel_value(self, field, val, center, h_ker):
        check(lib().shamb200_model_add_kernel_value(self.h, field.encode(), C.c_double(val),
                                                    (C.c_double * 3)(*center), C.c_double(h_ker)))

    def get_sum(self, field):
        o = (C.c_double * 3)()
        check(lib().shamb200_model_get_sum(self.h, field.encode(), o))
        return np.array(o[:])

    def total_part_count(self):
        n = C.c_uint64()
        check(lib().shamb200_model_total_part_count(self.h, C.byref(n)))
        return int(n.value)

    def set_particle_mass(self, gpart_mass):
        check(lib().shamb200_model_set_particle_mass(self.h, C.c_double(gpart_mass)))

    @property
    def patch_count(self):
        return lib().shamb200_model_patch_count(self.h)

    def patch_is_local(self, ip):
        return bool(lib().shamb200_model_patch_is_local(self.h, C.c_uint32(ip)))

    def patch_size(self, ip):
        return lib().shamb200_model_patch_size(self.h, C.c_uint32(ip))

    def get(self, ip, name):
        base = name.split(".")[-1]
        dt = np.uint32 if base in _U32 else (np.uint8 if base in _U8 else np.float64)
        nb = lib().shamb200_model_get(self.h, C.c_uint32(ip), name.encode(), None, C.c_int64(0))
        if nb == -1:
            raise KeyError(name)
        if nb < 0:
            raise ShamB200Error(lib().shamb200_last_error().decode())
        out = np.empty(nb // np.dtype(dt).itemsize, dtype=dt)
        if nb:
            r = lib().shamb200_model_get(self.h, C.c_uint32(ip), name.encode(), out.ctypes.data_as(C.c_void_p),
                                         C.c_int64(nb))
            if r < 0:
                raise ShamB200Error(lib().shamb200_last_error().decode())
        if base in _VEC3:
            out = out.reshape(-1, 3)
        return out

    def set_field(self, ip, name, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
        check(lib().shamb200_model_set_field(self.h, C.c_uint32(ip), name.encode(),
                                             a.ctypes.data_as(C.c_void_p), C.c_uint64(a.size)))

    def reorder_particles(self):
        """modules::ParticleReordering::reorder_particles: Morton order of every local patch"""
        check(lib().shamb200_model_reorder_particles(self.h))

    def evolve_once(self):
        check(lib().shamb200_model_evolve_once(self.h))
        return self.state()

    def evolve_once_host(self, ip, n, inputs, outputs):
        """Solver::evolve_once on host-resident patch data.  `inputs` / `outputs`: dict field name ->
        host address (int, page-locked memory) of n * nvar doubles; missing names are not copied."""
        hin, hout = HostPatchData(), HostPatchData()
        hin.n = hout.n = int(n)
        for h, d in ((hin, inputs), (hout, outputs)):
            for nm, addr in d.items():
                setattr(h, nm if nm != "uint" else "uint_", C.c_void_p(int(addr)))
        check(lib().shamb200_model_evolve_once_host(self.h, C.c_uint32(ip), C.byref(hin), C.byref(hout)))
        return int(hout.n)

    def search_stats(self):
        """(sum of the neighbour list lengths, accept tests done) of the last step on this rank"""
        o = (C.c_uint64 * 2)()
        check(lib().shamb200_model_search_stats(self.h, o))
        return int(o[0]), int(o[1])

    def conservation(self):
        """modules::ConservativeCheck of the last step: dict(sum_p, sum_a, sum_e, sum_de), all ranks"""
        o = (C.c_double * 8)()
        check(lib().shamb200_model_conservation(self.h, o))
        return dict(sum_p=np.array(o[0:3]), sum_a=np.array(o[3:6]), sum_e=float(o[6]), sum_de=float(o[7]))

    def host_traffic(self):
        o = (C.c_uint64 * 2)()
        check(lib().shamb200_model_host_traffic(self.h, o))
        return int(o[0]), int(o[1])

    def list_tolerance(self):
        """neighbour-list tolerance of the model's step: dict(last, growth, next, fallbacks)"""
        o = (C.c_double * 4)()
        check(lib().shamb200_model_list_tolerance(self.h, o))
        return dict(last=float(o[0]), growth=float(o[1]), next=float(o[2]), fallbacks=int(o[3]))

    def host_step_info(self, ip=0):
        """(id ranges the last host step was cut into, objects stored far from their Morton position)"""
        o = (C.c_uint64 * 2)()
        check(lib().shamb200_model_host_step_info(self.h, C.c_uint32(ip), o))
        return int(o[0]), int(o[1])

    def state(self):
        o = (C.c_double * 12)()
        check(lib().shamb200_model_state(self.h, o))
        keys = ("time", "dt", "cfl_multiplier", "eps_v", "h_subcycles", "h_iters_last", "corrector_iter",
                "npart", "t_step", "rate", "K_local", "n_local")
        return dict(zip(keys, list(o)))

    def set_next_dt(self, dt):
        check(lib().shamb200_model_set_next_dt(self.h, C.c_double(dt)))

    def set_time(self, t):
        check(lib().shamb200_model_set_time(self.h, C.c_double(t)))

    def set_cfl_multiplier(self, v):
        check(lib().shamb200_model_set_cfl_multiplier(self.h, C.c_double(v)))

    def stage_times(self):
        names = C.c_char_p()
        ms = C.POINTER(C.c_double)()
        cnt = C.c_uint32()
        check(lib().shamb200_model_stage_times(self.h, C.byref(names), C.byref(ms), C.byref(cnt)))
        ns = names.value.decode().split(";") if names.value else []
        return {n: ms[i] for i, n in enumerate(ns[: cnt.value])}


# ---- Phantom dump files (io_formats.cu; host only, no device needed) ----
def phantom_gen_config(fname, cfg=None, bypass_error=False):
    """Model::gen_config_from_phantom_dump: fills the fields the dump defines into cfg (default configuration if None)"""
    cfg = default_config() if cfg is None else cfg
    check(lib().shamb200_phantom_gen_config(str(fname).encode(), C.c_int(1 if bypass_error else 0), C.byref(cfg)))
    return cfg


def phantom_copy(fname_in, fname_out):
    check(lib().shamb200_phantom_copy(str(fname_in).encode(), str(fname_out).encode()))


def phantom_header_float(fname, key):
    out, found = C.c_double(0), C.c_int(0)
    check(lib().shamb200_phantom_header_float(str(fname).encode(), key.encode(), C.byref(out), C.byref(found)))
    return out.value if found.value else None


def phantom_header_int(fname, key):
    out, found = C.c_int64(0), C.c_int(0)
    check(lib().shamb200_phantom_header_int(str(fname).encode(), key.encode(), C.byref(out), C.byref(found)))
    return out.value if found.value else None


def phantom_compare(fname_a, fname_b):
    out = C.c_uint64(0)
    check(lib().shamb200_phantom_compare(str(fname_a).encode(), str(fname_b).encode(), C.byref(out)))
    return out.value


def plan_patch_grid(bmin, bmax, grid, world=1):
    """host-only: boxes [np,2,3] and owner rank [np] of the static patch grid"""
    n = grid[0] * grid[1] * grid[2]
    boxes = np.zeros((n, 2, 3))
    owner = np.zeros(n, dtype=np.int32)
    check(lib().shamb200_plan_patch_grid((C.c_double * 3)(*bmin), (C.c_double * 3)(*bmax), *[C.c_uint32(g) for g in grid],
                                         int(world), boxes.ctypes.data_as(C.c_void_p), owner.ctypes.data_as(C.c_void_p)))
    return boxes, owner


def plan_interfaces(boxes, bmin, bmax, periodic, interact_r, pcount):
    """host-only: the ghost interfaces in exchange order (list of Iface)"""
    boxes = np.ascontiguousarray(boxes, dtype=np.float64)
    n = len(boxes)
    ir = np.ascontiguousarray(interact_r, dtype=np.float64)
    pc = np.ascontiguousarray(pcount, dtype=np.uint32)
    cap = 27 * n * n
    out = (Iface * cap)()
    nf = C.c_uint32()
    check(lib().shamb200_plan_interfaces(C.c_uint32(n), boxes.ctypes.data_as(C.c_void_p), (C.c_double * 3)(*bmin),
                                         (C.c_double * 3)(*bmax), int(periodic), ir.ctypes.data_as(C.c_void_p),
                                         pc.ctypes.data_as(C.c_void_p), C.c_uint32(cap), out, C.byref(nf)))
    return [out[i] for i in range(nf.value)]


def nccl_unique_id():
    buf = (C.c_char * 128)()
    rc = lib().shamb200_nccl_unique_id(buf)
    if rc != 0:
        raise ShamB200Error("cannot create a NCCL unique id")
    return bytes(buf)
