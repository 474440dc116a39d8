"""Adaptive neighbour-list tolerance of the fast fp mode (solver.cu: Model::sph_prestep, shamb200_model_list_tolerance).

The reference builds the lists of a step with the radius R h htol, htol = 1.1 (Solver.cpp:1322-1386,
NeighbourCache.cpp:482-520).  The fast mode builds them with a tolerance fitted to the h growth of the previous step
and falls back to htol when some h outgrew it: the fields must agree with the oracle (which always uses htol) within
the 1e-10 contract either way, over several real timesteps.
"""
import numpy as np
import pytest

from tests import scenarios as S
from tests.test_gpu_step import close

pytestmark = pytest.mark.gpu

MAIN = ("xyz", "vxyz", "hpart", "uint", "axyz", "duint")


def fields_of(sc):
    names = list(MAIN)
    if sc["cfg"]["av"] in (2, 3):
        names += ["alpha_AV", "divv", "soundspeed"]
    if sc["cfg"]["av"] == 3:
        names += ["curlv", "dtdivv"]
    return names


def run(sc, steps, rtol=1e-10):
    o = S.make_oracle(sc)
    m = S.make_cuda(sc, fp_mode="fast", keep_step_data=False)
    tols, Ks = [], []
    for k in range(steps):
        so, sm = o.evolve_once(), m.evolve_once()
        assert so["npart"] == sm["npart"] and so["h_subcycles"] == sm["h_subcycles"], (k, so, sm)
        assert abs(sm["dt"] - so["dt"]) <= rtol * abs(so["dt"]), (k, so["dt"], sm["dt"])
        t = m.list_tolerance()
        tols.append(t)
        Ks.append(sm["K_local"])
        for ip in range(m.patch_count):
            assert m.patch_size(ip) == o.patch_size(ip)
            if not m.patch_size(ip):
                continue
            for nm in fields_of(sc):
                ok, msg = close(m.get(ip, nm), o.get(ip, nm), rtol)
                assert ok, f"step {k} patch {ip} {nm}: {msg} (list tolerance {t})"
    m.close()
    return tols, Ks


@pytest.mark.parametrize("scenario", ["periodic_cd10", "periodic_m6", "multi_patch", "disc", "sod"])
def test_adaptive_list_tolerance_matches_oracle(scenario):
    sc = {
        "periodic_cd10": lambda: S.periodic_box(6000, "M4", "cd10", jitter=0.15),
        "periodic_m6": lambda: S.periodic_box(5000, "M6", "mm97", jitter=0.1),
        "multi_patch": lambda: S.periodic_box(9000, "M4", "cd10", jitter=0.1, grid=(2, 2, 1)),
        "disc": lambda: S.disc(3000, "M4"),
        "sod": lambda: S.sod_tube(resol=16, kernel="M4"),
    }[scenario]()
    tols, Ks = run(sc, steps=5)
    assert tols[0]["last"] == pytest.approx(1.1)  # no history: the reference's tolerance
    # later steps use what the growth of the step before asked for (or fell back to 1.1 when that was not enough)
    for prev, cur in zip(tols, tols[1:]):
        assert cur["last"] in (pytest.approx(prev["next"]), pytest.approx(1.1)), (prev, cur)
        assert cur["growth"] <= cur["last"]
    assert min(t["last"] for t in tols[1:]) < 1.1
    assert min(Ks[1:]) < Ks[0]


def test_list_tolerance_fallback(monkeypatch):
    """a starting tolerance far too small for the h growth of a jittered box: every step redoes its sub-cycle with
    the reference's tolerance and still matches the oracle"""
    monkeypatch.setenv("SHAMB200_LIST_TOL", "1.0000001")
    sc = S.periodic_box(6000, "M4", "cd10", jitter=0.15)
    tols, _ = run(sc, steps=3)
    assert tols[-1]["fallbacks"] >= 2
    assert all(t["last"] == pytest.approx(1.1) for t in tols)


def test_list_tolerance_off(monkeypatch):
    monkeypatch.setenv("SHAMB200_LIST_TOL", "0")
    sc = S.periodic_box(6000, "M4", "cd10", jitter=0.15)
    tols, Ks = run(sc, steps=3)
    assert all(t["last"] == pytest.approx(1.1) for t in tols) and tols[-1]["fallbacks"] == 0


def test_tight_lists_equal_full_lists_on_replay():
    """dt = 0 replays (the bench protocol): h does not move, the lists shrink to the 0.5 % floor and every field
    equals the full-list result to round-off (the entries left out contribute exactly zero)"""
    sc = S.periodic_box(20000, "M4", "cd10", jitter=0.05)
    import os
    os.environ["SHAMB200_LIST_TOL"] = "0"
    try:
        full = S.make_cuda(sc, fp_mode="fast", keep_step_data=False)
        full.evolve_once()
        for _ in range(2):
            full.set_next_dt(0.0)
            full.evolve_once()
    finally:
        del os.environ["SHAMB200_LIST_TOL"]
    m = S.make_cuda(sc, fp_mode="fast", keep_step_data=False)
    m.evolve_once()
    for _ in range(2):
        m.set_next_dt(0.0)
        m.evolve_once()
    t = m.list_tolerance()
    assert t["last"] == pytest.approx(1.005) and t["fallbacks"] == 0
    assert m.state()["K_local"] < 0.85 * full.state()["K_local"]
    for nm in ("xyz", "vxyz", "hpart", "uint", "axyz", "duint", "alpha_AV", "divv", "curlv", "dtdivv", "soundspeed"):
        ok, msg = close(m.get(0, nm), full.get(0, nm), 1e-12)
        assert ok, f"{nm}: {msg}"
    assert m.state()["dt"] == pytest.approx(full.state()["dt"], rel=1e-12)
