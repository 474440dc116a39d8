"""CPU-side checks of the drop-in boundary: libshamb200.so loads without a GPU, exports every symbol
include/shamb200.h declares, the ctypes mirror of shamb200_solver_config has the C layout, and the
compute entry points fail loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from shamrock_b200 import _capi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _capi.lib()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "shamb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(shamb200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/shamb200.h but not exported"
    assert sorted(_capi.SYMBOLS) == names  # the python mirror lists the same surface


def test_build_info_and_launch_counter(lib):
    info = lib.shamb200_build_info().decode()
    assert "sm_100a" in info
    lib.shamb200_reset_launch_count()
    assert lib.shamb200_launch_count() == 0


def test_default_config_matches_reference_defaults(lib):
    cfg = _capi.default_config()
    # SolverConfig.hpp:584-610, AVConfig.hpp, Solver.hpp (cfl) defaults
    assert cfg.tree_reduction_level == 3 and cfg.use_two_stage_search == 1
    assert cfg.htol_up_coarse_cycle == 1.1 and cfg.htol_up_fine_cycle == 1.1 and cfg.epsilon_h == 1e-6
    assert cfg.h_iter_per_subcycles == 50 and cfg.h_max_subcycles_count == 100
    assert cfg.cfl_multiplier_stiffness == 2 and cfg.gamma == 5.0 / 3.0
    # layout: the last field written by the C side is where ctypes thinks it is
    assert C.sizeof(_capi.SolverConfig) % 8 == 0 and cfg.constant_G == 1.0


def test_default_config_reordering_fields(lib):
    cfg = _capi.default_config()  # SolverConfig.hpp:621-630; the new fields sit at the end of the C struct
    assert cfg.enable_particle_reordering == 0 and cfg.particle_reordering_step_freq == 1000
    assert cfg.n_kill_spheres == 0 and cfg.kill_radius[3] == 0.0


def test_stale_handles_are_refused_not_dereferenced(lib):
    """handles that are not live (never created, or destroyed): destroy is a no-op, use is SHAMB200_ERR_INVALID"""
    fake_m, fake_c = C.c_void_p(0x1000), C.c_void_p(0x2000)
    assert lib.shamb200_model_destroy(fake_m) == 0
    assert lib.shamb200_ctx_destroy(fake_c) == 0
    assert lib.shamb200_model_patch_count(fake_m) == 0 and lib.shamb200_model_patch_size(fake_m, C.c_uint32(0)) == 0
    assert lib.shamb200_model_evolve_once(fake_m) == -1
    assert b"stale model handle" in lib.shamb200_last_error()
    assert lib.shamb200_model_set_next_dt(fake_m, C.c_double(0.0)) == -1
    assert lib.shamb200_model_reorder_particles(fake_m) == -1
    assert lib.shamb200_ctx_synchronize(fake_c) == -1
    assert b"stale context handle" in lib.shamb200_last_error()
    assert lib.shamb200_ctx_stream(fake_c) is None
    out = C.c_void_p()
    assert lib.shamb200_model_create(fake_c, C.byref(_capi.default_config()), C.byref(out)) == -1
    d = C.c_double()
    assert lib.shamb200_microbench(fake_c, 0, C.byref(d)) == -1
    z = (C.c_double * 3)()
    assert lib.shamb200_leapfrog_predict(fake_c, C.c_uint32(0), C.c_double(0.1), z, z, z, z, z) == -1


def test_no_cpu_fallback_without_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(_capi.ShamB200Error, match="no CPU fallback"):
        _capi.Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "shamrock_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                for pat in ("pyoracle", "liboracle", "import oracle", "from oracle", "#include \"../../oracle"):
                    assert pat not in txt, f"{f} references the oracle ({pat})"
