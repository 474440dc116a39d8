"""Patch -> rank load balancing along the Hilbert curve (HilbertLoadBalance + load_balance of
the reference): the host-side planner of the C ABI (shamb200_hilbert_index, shamb200_plan_load_balance; no CUDA
call) against the oracle restatement and the reference's own known answer."""
import numpy as np
import pytest

from oracle import load_balance as olb
from shamrock_b200 import _capi

GRID = 1 << 21


def test_hilbert_known_answer_of_the_reference():
    # src/tests/shamrock/patch/legacy/scheduler/test_hilbert_sfc.cpp
    assert olb.compute_hilbert_index_3d(GRID - 1, 0, 0) == 9223372036854775807
    assert _capi.hilbert_index(GRID - 1, 0, 0) == 9223372036854775807
    assert _capi.hilbert_index(0, 0, 0) == 0 == olb.compute_hilbert_index_3d(0, 0, 0)


def test_hilbert_index_matches_oracle_and_is_a_curve():
    rng = np.random.default_rng(5)
    pts = rng.integers(0, GRID, size=(300, 3), dtype=np.uint64)
    for x, y, z in pts:
        assert _capi.hilbert_index(x, y, z) == olb.compute_hilbert_index_3d(int(x), int(y), int(z))
    # the cells of a 16^3 patch grid: distinct indices, consecutive ones are face neighbours
    n = 16
    cell = GRID // n
    cells = [(x, y, z) for x in range(n) for y in range(n) for z in range(n)]
    codes = [_capi.hilbert_index(x * cell, y * cell, z * cell) for x, y, z in cells]
    assert len(set(codes)) == n**3
    order = np.argsort(np.array(codes, dtype=np.uint64))
    walk = np.array(cells)[order]
    assert np.all(np.abs(np.diff(walk, axis=0)).sum(axis=1) == 1)


@pytest.mark.parametrize("world", [1, 2, 3, 8, 64])
@pytest.mark.parametrize("kind", ["uniform", "disc", "empty", "one_heavy"])
def test_plan_load_balance_matches_oracle(world, kind):
    rng = np.random.default_rng(world * 7 + len(kind))
    n = 8
    cell = GRID // n
    coords = np.array([(x * cell, y * cell, z * cell) for z in range(n) for y in range(n) for x in range(n)],
                      dtype=np.uint64)
    if kind == "uniform":  # the reference's test: loads within 20 % of each other
        load = rng.integers(1000000, 1200000, size=len(coords))
    elif kind == "disc":  # most patches empty, the load in a thin slab
        c = coords.astype(np.float64) / GRID - 0.5 + 0.5 / n
        r = np.hypot(c[:, 0], c[:, 1])
        load = np.where((np.abs(c[:, 2]) < 0.1) & (r > 0.15) & (r < 0.45), rng.integers(1000, 90000, len(coords)), 0)
    elif kind == "empty":
        load = np.zeros(len(coords), dtype=np.int64)
    else:
        load = np.ones(len(coords), dtype=np.int64)
        load[37] = 10**9
    owner, strat = _capi.plan_load_balance(coords, load, world)
    ref_owner, ref_strat = olb.hilbert_load_balance(coords, load, world)
    assert strat == ref_strat
    assert np.array_equal(owner, np.array(ref_owner, dtype=np.int32))
    assert owner.min() >= 0 and owner.max() < world
    # a sweep: owners never decrease along the curve
    codes = np.array([_capi.hilbert_index(*c) for c in coords], dtype=np.uint64)
    assert np.all(np.diff(owner[np.argsort(codes)]) >= 0)
    if kind == "uniform" and world <= 64:
        per = np.bincount(owner, weights=load, minlength=world)
        assert per.max() <= 1.35 * load.sum() / world  # 512 tiles on <= 64 ranks: within a tile or two of even


def test_plan_load_balance_rejects_bad_arguments():
    with pytest.raises(_capi.ShamB200Error):
        _capi.plan_load_balance(np.zeros((2, 3), dtype=np.uint64), np.ones(2, dtype=np.uint64), 0)


@pytest.mark.gpu
def test_model_patch_owner_table():
    pytest.importorskip("torch")
    from tests import scenarios as S

    sc = S.periodic_box(4000, "M4", "cd10", jitter=0.1, grid=(2, 2, 2))
    ctx = _capi.Context(0)
    cfg = _capi.default_config()
    m = _capi.Model(ctx, cfg)
    m.set_box(sc["bmin"], sc["bmax"], sc["grid"])
    coords = m.patch_coords()
    assert coords.shape == (8, 3) and coords.max() == GRID // 2
    owner, _ = _capi.plan_load_balance(coords, np.ones(8), 1)
    m.set_patch_owners(owner)  # world size 1: everything stays local
    with pytest.raises(_capi.ShamB200Error):
        m.set_patch_owners(np.ones(8, dtype=np.int32))  # rank 1 does not exist
    m.push_particles(sc["xyz"], sc["vxyz"], sc["hpart"], sc["uint"])
    assert sum(m.patch_size(ip) for ip in range(8)) == len(sc["xyz"])
    with pytest.raises(_capi.ShamB200Error):
        m.set_patch_owners(owner)  # particles are in place
    m.close()
    ctx.close()
