"""modules::ParticleReordering (shammodels/sph/src/modules/ParticleReordering.cpp:22-51):
every patch permuted into the Morton order of its positions over the patch box, all fields together.

CPU: the oracle's restatement (sorted codes, bijection, every field follows its particle, idempotence).
GPU: shamb200_model_reorder_particles and the in-step reordering (enable_particle_reordering /
particle_reordering_step_freq, Solver.cpp:2043-2048) against the oracle, bit for bit; and the physics does not
depend on the storage order beyond the summation order (1e-10)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests import scenarios as S

FIELDS = ["xyz", "vxyz", "axyz", "axyz_ext", "hpart", "uint", "duint"]


def tagged(sc):
    """a scenario whose uint field numbers the particles (the tag follows the particle)"""
    sc = dict(sc)
    sc["uint"] = np.arange(len(sc["xyz"]), dtype=np.float64) + 1.0
    sc["vxyz"] = np.stack([sc["uint"] * 2, sc["uint"] * 3, sc["uint"] * 5], axis=1)
    return sc


def patch_boxes(sc):
    """[lo, hi) of the patches of the static grid, patch id = x + nx (y + ny z) (oracle/sph_step.hpp patch_box)"""
    nx, ny, nz = sc["grid"]
    bmin, bmax = np.array(sc["bmin"], dtype=np.float64), np.array(sc["bmax"], dtype=np.float64)
    G = float(1 << 21)
    out = []
    for z in range(nz):
        for y in range(ny):
            for x in range(nx):
                fact = (bmax - bmin) / G
                sz = np.array([(1 << 21) // nx, (1 << 21) // ny, (1 << 21) // nz], dtype=np.float64)
                c = np.array([x, y, z], dtype=np.float64)
                out.append((sz * c * fact + bmin, sz * (c + 1) * fact + bmin))
    return out


@pytest.mark.parametrize("grid", [(1, 1, 1), (2, 2, 1)])
def test_oracle_reorder_particles(grid):
    sc = tagged(S.periodic_box(3000, "M4", "cd10", jitter=0.3, grid=grid))
    o = S.make_oracle(sc)
    before = [{k: o.get(ip, k).copy() for k in FIELDS} for ip in range(o.patch_count)]
    o.reorder_particles()
    boxes = patch_boxes(sc)
    for ip in range(o.patch_count):
        n = o.patch_size(ip)
        assert n == len(before[ip]["uint"]) and n > 0
        xyz, tag = o.get(ip, "xyz"), o.get(ip, "uint")
        lo, hi = boxes[ip]
        codes = po.morton_codes(xyz, lo, hi, n)
        assert np.all(np.diff(codes.astype(np.int64)) >= 0), "positions are not in Morton order"
        # a permutation of the patch, and every field moved with its particle
        old_tag = before[ip]["uint"]
        assert np.array_equal(np.sort(tag), np.sort(old_tag))
        src = {t: i for i, t in enumerate(old_tag)}
        perm = np.array([src[t] for t in tag])
        for k in FIELDS:
            assert np.array_equal(o.get(ip, k), before[ip][k][perm]), k
        # ties keep the order of the reference's bitonic network on (code, index) pairs padded to a power of two
        P2 = 1 << max(int(n - 1).bit_length(), 0)
        c0 = po.morton_codes(before[ip]["xyz"], lo, hi, P2)
        _, ids = po.sort_by_key(c0)
        assert np.array_equal(ids[:n], perm)
    after = [{k: o.get(ip, k).copy() for k in FIELDS} for ip in range(o.patch_count)]
    o.reorder_particles()  # sorted input: equal codes may still swap inside the network, positions stay sorted
    for ip in range(o.patch_count):
        lo, hi = boxes[ip]
        n = o.patch_size(ip)
        assert np.array_equal(po.morton_codes(o.get(ip, "xyz"), lo, hi, n), po.morton_codes(after[ip]["xyz"], lo, hi, n))


def test_oracle_step_with_reordering_matches_without():
    """the step does not depend on the storage order beyond the order of the sums"""
    sc = S.periodic_box(2500, "M4", "cd10", jitter=0.2)
    sc_r = dict(sc, cfg=dict(sc["cfg"], enable_particle_reordering=1, particle_reordering_step_freq=1))
    a, b = S.make_oracle(sc), S.make_oracle(sc_r)
    for _ in range(2):
        sa, sb = a.evolve_once(), b.evolve_once()
    assert abs(sa["dt"] - sb["dt"]) <= 1e-12 * sa["dt"]
    # match the particles by position (the drift is order independent up to rounding of the forces)
    xa, xb = a.get(0, "xyz"), b.get(0, "xyz")
    ia, ib = np.lexsort(np.round(xa, 9).T), np.lexsort(np.round(xb, 9).T)
    assert np.allclose(xa[ia], xb[ib], rtol=0, atol=1e-12)
    for k in ("hpart", "uint", "duint"):
        va, vb = a.get(0, k)[ia], b.get(0, k)[ib]
        assert np.allclose(va, vb, rtol=1e-10, atol=1e-10 * np.abs(va).mean()), k


# ---- GPU --------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("grid,jitter", [((1, 1, 1), 0.3), ((2, 2, 1), 0.3), ((1, 1, 1), 0.0)])
def test_gpu_reorder_particles_matches_oracle(grid, jitter):
    pytest.importorskip("torch")
    sc = tagged(S.periodic_box(5000, "M4", "cd10", jitter=jitter, grid=grid))
    o, m = S.make_oracle(sc), S.make_cuda(sc)
    o.reorder_particles()
    m.reorder_particles()
    for ip in range(o.patch_count):
        assert m.patch_size(ip) == o.patch_size(ip)
        for k in FIELDS:
            assert np.array_equal(m.get(ip, k), o.get(ip, k)), (ip, k)
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fp_mode", ["strict", "fast"])
def test_gpu_step_with_reordering(fp_mode):
    pytest.importorskip("torch")
    from tests.test_gpu_step import run_and_compare

    sc = S.periodic_box(6000, "M4", "cd10", jitter=0.15, grid=(2, 1, 1))
    sc["cfg"] = dict(sc["cfg"], enable_particle_reordering=1, particle_reordering_step_freq=2)
    run_and_compare(sc, steps=3, fp_mode=fp_mode)  # reorders at steps 0 and 2


@pytest.mark.gpu
def test_gpu_radix_reorder_is_a_sorted_permutation():
    """sort_mode radix (the bench mode): same sorted codes, ties in input order"""
    pytest.importorskip("torch")
    sc = tagged(S.periodic_box(20000, "M4", "cd10", jitter=0.0, sort_mode="radix"))
    m = S.make_cuda(sc)
    m.reorder_particles()
    xyz, tag = m.get(0, "xyz"), m.get(0, "uint")
    lo, hi = patch_boxes(sc)[0]
    codes = po.morton_codes(xyz, lo, hi, len(xyz)).astype(np.int64)
    assert np.all(np.diff(codes) >= 0)
    assert np.array_equal(np.sort(tag), sc["uint"])
    same = np.diff(codes) == 0
    assert np.all(np.diff(tag)[same] > 0), "a stable sort keeps equal codes in input order"
    assert np.array_equal(xyz, sc["xyz"][(tag - 1).astype(np.int64)])
    m.close()
