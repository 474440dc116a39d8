"""ctypes front of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under shamrock_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")


_STAMP = os.path.join(_HERE, "_build", "cpu.stamp")


def _cpu_id():
    """model + ISA flags of this host: the library is built with -march=native"""
    try:
        model, flags = "", ""
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name") and not model:
                model = line.split(":", 1)[1].strip()
            if line.startswith("flags") and not flags:
                flags = " ".join(sorted(line.split(":", 1)[1].split()))
            if model and flags:
                break
        import hashlib

        return model + " " + hashlib.sha1(flags.encode()).hexdigest()
    except OSError:
        return "unknown"


def build(force=False):
    """Compile the oracle (g++, OpenMP) if needed: sources newer than the library, or a library that was built
    on another CPU (-march=native; the in-tree .so travels to the GPU box)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "shamrock_oracle.hpp", "sph_step.hpp", "Makefile")]
    same_cpu = os.path.exists(_STAMP) and open(_STAMP).read() == _cpu_id()
    if (not force) and same_cpu and os.path.exists(_LIB) and all(
        os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in srcs
    ):
        return _LIB
    subprocess.check_call(["make", "-C", _HERE, "-B", "-s"])
    with open(_STAMP, "w") as f:
        f.write(_cpu_id())
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_tree_build.restype = C.c_void_p
        L.oracle_tree_get.restype = C.c_int64
        L.oracle_solver_create.restype = C.c_void_p
        L.oracle_solver_get.restype = C.c_int64
        L.oracle_kernel_eval.restype = C.c_double
        L.oracle_solver_patch_count.restype = C.c_uint32
        L.oracle_solver_patch_size.restype = C.c_uint32
        _lib = L
    return _lib


def _chk(rc):
    if rc != 0:
        raise RuntimeError(lib().oracle_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


KERNELS = {"M4": 0, "M6": 1}
_MT = {32: np.uint32, 64: np.uint64}


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


def morton_codes(xyz, bmin, bmax, morton_count, bits=32):
    xyz = _f64(xyz).reshape(-1, 3)
    out = np.empty(morton_count, dtype=_MT[bits])
    _chk(lib().oracle_morton_codes(bits, _p(xyz), C.c_uint64(3), C.c_uint32(len(xyz)), _p(_f64(bmin)),
                                   _p(_f64(bmax)), C.c_uint32(morton_count), _p(out)))
    return out


def sort_by_key(keys, vals=None, bits=32):
    keys = np.array(keys, dtype=_MT[bits])
    vals = np.arange(len(keys), dtype=np.uint32) if vals is None else np.array(vals, dtype=np.uint32)
    _chk(lib().oracle_sort_by_key(bits, _p(keys), _p(vals), C.c_uint32(len(keys))))
    return keys, vals


def reduction(sorted_codes, morton_count, level, bits=32):
    sorted_codes = np.ascontiguousarray(sorted_codes, dtype=_MT[bits])
    out = np.zeros(morton_count + 2, dtype=np.uint32)
    lc = C.c_uint32(0)
    _chk(lib().oracle_reduction(bits, _p(sorted_codes), C.c_uint32(morton_count), C.c_uint32(level), _p(out),
                                C.byref(lc)))
    return out[: lc.value + 2].copy(), lc.value


def karras(codes, bits=32):
    codes = np.ascontiguousarray(codes, dtype=_MT[bits])
    n = max(len(codes) - 1, 0)
    lc, rc, er = (np.zeros(n, dtype=np.uint32) for _ in range(3))
    lf, rf = (np.zeros(n, dtype=np.uint8) for _ in range(2))
    _chk(lib().oracle_karras(bits, _p(codes), C.c_uint32(len(codes)), _p(lc), _p(rc), _p(lf), _p(rf), _p(er)))
    return dict(lchild_id=lc, rchild_id=rc, lchild_flag=lf, rchild_flag=rf, endrange=er)


_TREE_DT = {
    "sort_index_map": np.uint32, "reduc_index_map": np.uint32, "lchild_id": np.uint32,
    "rchild_id": np.uint32, "lchild_flag": np.uint8, "rchild_flag": np.uint8, "endrange": np.uint32,
    "aabb_min": np.float64, "aabb_max": np.float64, "rint": np.float64, "leaf_owner": np.uint32,
    "cnt_neigh": np.uint32, "scanned_cnt": np.uint32, "index_neigh_map": np.uint32,
}


class Tree:
    """CompressedLeafBVH built by the oracle (shamtree::CompressedLeafBVH::rebuild_from_positions)."""

    def __init__(self, xyz, bmin, bmax, level, bits=32, morton_count=0):
        self.bits = bits
        self.xyz = _f64(xyz).reshape(-1, 3)
        self.h = lib().oracle_tree_build(bits, _p(self.xyz), C.c_uint64(3), C.c_uint32(len(self.xyz)),
                                         _p(_f64(bmin)), _p(_f64(bmax)), C.c_uint32(level), C.c_uint32(morton_count))
        if not self.h:
            raise RuntimeError(lib().oracle_last_error().decode())
        s = (C.c_uint32 * 4)()
        lib().oracle_tree_sizes(C.c_void_p(self.h), s)
        self.obj_cnt, self.morton_count, self.leaf_count, self.int_count = list(s)

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_tree_free(C.c_void_p(self.h))
            self.h = None

    def get(self, name):
        base = name.split(".")[-1]
        dt = _MT[self.bits] if base in ("sorted_morton", "reduced_morton") else _TREE_DT[base]
        nb = lib().oracle_tree_get(C.c_void_p(self.h), name.encode(), None, C.c_int64(0))
        if nb < 0:
            raise KeyError(name)
        out = np.empty(nb // np.dtype(dt).itemsize, dtype=dt)
        lib().oracle_tree_get(C.c_void_p(self.h), name.encode(), _p(out), C.c_int64(nb))
        if base in ("aabb_min", "aabb_max"):
            out = out.reshape(-1, 3)
        return out

    def field_max(self, field, htol=1.0):
        _chk(lib().oracle_tree_field_max(C.c_void_p(self.h), _p(_f64(field)), C.c_double(htol)))
        return self.get("rint")

    def neigh_cache(self, hpart, obj_cnt, Rkern, htol, two_stage=True, xyz=None):
        x = self.xyz if xyz is None else _f64(xyz).reshape(-1, 3)
        _chk(lib().oracle_tree_neigh_cache(C.c_void_p(self.h), _p(x), C.c_uint64(3), _p(_f64(hpart)),
                                           C.c_uint32(obj_cnt), C.c_double(Rkern), C.c_double(htol),
                                           int(two_stage)))
        return {k: self.get("cache." + k) for k in ("cnt_neigh", "scanned_cnt", "index_neigh_map")}

    def box_query(self, s):
        _chk(lib().oracle_tree_box_query(C.c_void_p(self.h), _p(self.xyz), C.c_uint64(3),
                                         C.c_uint32(len(self.xyz)), C.c_double(s)))
        return {k: self.get("cache." + k) for k in ("cnt_neigh", "scanned_cnt", "index_neigh_map")}


def h_iterate(kernel, cache, xyz, h_old, h_new, eps, pmass, h_evol_max, h_evol_iter_max):
    """One sweep of IterateSmoothingLengthDensity; h_new and eps are updated in place."""
    xyz = _f64(xyz).reshape(-1, 3)
    cnt = np.ascontiguousarray(cache["cnt_neigh"], dtype=np.uint32)
    sc = np.ascontiguousarray(cache["scanned_cnt"], dtype=np.uint32)
    idx = np.ascontiguousarray(cache["index_neigh_map"], dtype=np.uint32)
    assert h_new.dtype == np.float64 and eps.dtype == np.float64
    _chk(lib().oracle_h_iterate(KERNELS[kernel], _p(cnt), _p(sc), _p(idx), C.c_uint64(len(idx)), _p(xyz),
                                C.c_uint64(3), C.c_uint32(len(cnt)), _p(_f64(h_old)), _p(h_new), _p(eps),
                                C.c_double(pmass), C.c_double(h_evol_max), C.c_double(h_evol_iter_max)))


def kernel_eval(kernel, which, a, b=1.0):
    w = {"f": 0, "df": 1, "W_3d": 2, "dW_3d": 3, "dhW_3d": 4}[which]
    return lib().oracle_kernel_eval(KERNELS[kernel], w, C.c_double(a), C.c_double(b))


_STEP_DT = {"tree.sorted_morton": np.uint32, "tree.reduced_morton": np.uint32}


class Solver:
    """Oracle twin of shammodels::sph::Solver (evolve_once on CPU)."""

    def __init__(self, cfg, bmin, bmax, patch_grid=(1, 1, 1)):
        self.s = lib().oracle_solver_create()
        self.configure(cfg)
        _chk(lib().oracle_solver_set_box(C.c_void_p(self.s), _p(_f64(bmin)), _p(_f64(bmax)),
                                         *[C.c_uint32(v) for v in patch_grid]))

    def __del__(self):
        if getattr(self, "s", None):
            lib().oracle_solver_free(C.c_void_p(self.s))
            self.s = None

    def configure(self, cfg):
        kv = ";".join(f"{k}={float(v)!r}" for k, v in cfg.items())
        _chk(lib().oracle_solver_configure(C.c_void_p(self.s), kv.encode()))

    def add_kill_sphere(self, center, radius):
        _chk(lib().oracle_solver_add_kill_sphere(C.c_void_p(self.s), _p(_f64(center)), C.c_double(radius)))

    def push_particles(self, xyz, vxyz, h, u):
        xyz = _f64(xyz).reshape(-1, 3)
        v = None if vxyz is None else _f64(vxyz).reshape(-1, 3)
        uu = None if u is None else _f64(u)
        _chk(lib().oracle_solver_push_particles(C.c_void_p(self.s), C.c_uint32(len(xyz)), _p(xyz),
                                                _p(v) if v is not None else None, _p(_f64(h)),
                                                _p(uu) if uu is not None else None))

    @property
    def patch_count(self):
        return lib().oracle_solver_patch_count(C.c_void_p(self.s))

    def patch_size(self, ip):
        return lib().oracle_solver_patch_size(C.c_void_p(self.s), C.c_uint32(ip))

    def set_field(self, ip, name, arr):
        _chk(lib().oracle_solver_set_field(C.c_void_p(self.s), C.c_uint32(ip), name.encode(), _p(_f64(arr))))

    def get(self, ip, name):
        base = name.split(".")[-1]
        if name in _STEP_DT:
            dt = _STEP_DT[name]
        elif name.startswith("tree.") or name.startswith("cache."):
            dt = _TREE_DT[base]
        else:
            dt = np.float64
        nb = lib().oracle_solver_get(C.c_void_p(self.s), C.c_uint32(ip), name.encode(), None, C.c_int64(0))
        if nb < 0:
            raise KeyError(name)
        out = np.empty(nb // np.dtype(dt).itemsize, dtype=dt)
        lib().oracle_solver_get(C.c_void_p(self.s), C.c_uint32(ip), name.encode(), _p(out), C.c_int64(nb))
        if base in ("xyz", "vxyz", "axyz", "axyz_ext", "curlv", "mxyz", "g_v", "g_a", "aabb_min", "aabb_max"):
            out = out.reshape(-1, 3)
        return out

    def reorder_particles(self):
        _chk(lib().oracle_solver_reorder_particles(C.c_void_p(self.s)))

    def split_patch(self, ip):
        _chk(lib().oracle_solver_split_patch(C.c_void_p(self.s), C.c_uint32(ip)))

    def merge_patches(self, ip0):
        _chk(lib().oracle_solver_merge_patches(C.c_void_p(self.s), C.c_uint32(ip0)))

    def patch_id(self, ip):
        lib().oracle_solver_patch_id.restype = C.c_uint64
        return int(lib().oracle_solver_patch_id(C.c_void_p(self.s), C.c_uint32(ip)))

    def evolve_once(self):
        _chk(lib().oracle_solver_evolve_once(C.c_void_p(self.s)))
        return self.state()

    def state(self):
        o = (C.c_double * 8)()
        lib().oracle_solver_state(C.c_void_p(self.s), o)
        keys = ("time", "dt", "cfl_multiplier", "eps_v", "h_subcycles", "h_iters_last", "corrector_iter", "npart")
        return dict(zip(keys, list(o)))

    def set_next_dt(self, dt):
        lib().oracle_solver_set_dt(C.c_void_p(self.s), C.c_double(dt))

    def set_particle_mass(self, m):
        lib().oracle_solver_set_gpart_mass(C.c_void_p(self.s), C.c_double(m))
