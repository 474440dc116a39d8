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
