"""Initial conditions generated on the device (shamrock_b200/csrc/setup.cu) against the host
restatement of the reference's generators (shamrock_b200/lattice.py = shammath::LatticeHCP,
crystalLattice.hpp:52-290) and numpy versions of Model::set_value_in_a_box / add_kernel_value / get_sum
(Model.hpp:669-785)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from shamrock_b200 import _capi, lattice  # noqa: E402
from tests import scenarios as S  # noqa: E402


def empty_model(sc, grid=None):
    sc = dict(sc)
    if grid:
        sc["grid"] = grid
    for k in ("xyz", "vxyz", "hpart", "uint"):
        sc[k] = sc[k][:0]
    return S.make_cuda(sc), sc


@pytest.mark.parametrize("grid", [(1, 1, 1), (2, 2, 1), (4, 1, 2)])
def test_device_lattice_is_bit_identical_to_the_host_generator(grid):
    """positions, order inside every patch, hpart = dr, all other fields 0 — and the same step afterwards"""
    ref_sc = S.periodic_box(9000, "M4", "cd10", jitter=0.0, grid=grid, inject=False)
    host = S.make_cuda(ref_sc)
    dev, sc = empty_model(ref_sc)
    dr = ref_sc["dr"]
    n = dev.add_lattice_hcp(dr, ref_sc["bmin"], ref_sc["bmax"])
    assert n == len(ref_sc["xyz"]) == dev.total_part_count()
    for ip in range(host.patch_count):
        assert dev.patch_size(ip) == host.patch_size(ip)
        assert np.array_equal(dev.get(ip, "xyz"), host.get(ip, "xyz")), f"patch {ip}"
        assert np.array_equal(dev.get(ip, "hpart"), np.full(dev.patch_size(ip), dr))
        for nm in ("vxyz", "axyz", "uint", "duint", "alpha_AV", "soundspeed"):
            assert not dev.get(ip, nm).any(), nm
    # the reference's setters on the device = the scenario's numpy setup
    h0 = ref_sc["hpart"][0]
    big = ([-1e30] * 3, [1e30] * 3)
    dev.set_value_in_a_box("hpart", h0, *big)
    dev.set_value_in_a_box("uint", 1.0, *big)
    for ip in range(host.patch_count):
        assert np.array_equal(dev.get(ip, "hpart"), host.get(ip, "hpart"))
        assert np.array_equal(dev.get(ip, "uint"), host.get(ip, "uint"))
    a, b = dev.evolve_once(), host.evolve_once()
    assert a["dt"] == b["dt"] and a["npart"] == b["npart"]
    for ip in range(host.patch_count):
        for nm in ("xyz", "axyz", "duint", "hpart"):
            assert np.array_equal(dev.get(ip, nm), host.get(ip, nm)), (ip, nm)


def test_sub_box_and_two_lattices():
    """Sod-like setup: two lattices of different spacing side by side (SPHSetup combiner_add), box setters"""
    sc = S.sod_tube(12, "M6", grid=(2, 1, 1))
    dev, _ = empty_model(sc)
    dr = sc["dr"]
    fact = 2.0
    bmin, bmax = sc["bmin"], sc["bmax"]
    n1 = dev.add_lattice_hcp(dr, bmin, (0.0, bmax[1], bmax[2]))
    n2 = dev.add_lattice_hcp(dr * fact, (0.0, bmin[1], bmin[2]), bmax)
    left = lattice.hcp_positions(dr, bmin, (0.0, bmax[1], bmax[2]))
    right = lattice.hcp_positions(dr * fact, (0.0, bmin[1], bmin[2]), bmax)
    assert (n1, n2) == (len(left), len(right))
    got = np.concatenate([dev.get(ip, "xyz") for ip in range(dev.patch_count)])
    want = np.concatenate([left, right])
    assert sorted(map(tuple, got)) == sorted(map(tuple, want))
    dev.set_value_in_a_box("uint", 2.5, bmin, (0.0, bmax[1], bmax[2]))
    dev.set_value_in_a_box("vxyz", [1.0, 2.0, 3.0], (0.0, bmin[1], bmin[2]), bmax)
    u = np.concatenate([dev.get(ip, "uint") for ip in range(dev.patch_count)])
    v = np.concatenate([dev.get(ip, "vxyz") for ip in range(dev.patch_count)])
    assert np.array_equal(u, np.where(got[:, 0] < 0.0, 2.5, 0.0))
    assert np.array_equal(v, np.where(got[:, :1] >= 0.0, np.array([[1.0, 2.0, 3.0]]), 0.0))
    s = dev.get_sum("uint")
    assert abs(s[0] - u.sum()) <= 1e-12 * abs(u.sum())
    sv = dev.get_sum("vxyz")
    assert np.allclose(sv, v.sum(0), rtol=1e-12)


@pytest.mark.parametrize("kernel", ["M4", "M6"])
def test_add_kernel_value_and_sphere(kernel):
    sc = S.periodic_box(5000, kernel, "cd10", jitter=0.1, inject=False)
    m = S.make_cuda(sc)
    x = m.get(0, "xyz")
    c, hk = (0.05, -0.02, 0.1), 6 * sc["dr"]
    m.add_kernel_value("uint", 3.0, c, hk)
    r = np.linalg.norm(x - np.array(c), axis=1)
    q = r / hk
    if kernel == "M4":
        f = np.where(q < 1, 0.25 * (2 - q) ** 3 - (1 - q) ** 3, np.where(q < 2, 0.25 * (2 - q) ** 3, 0.0)) / math.pi
    else:
        t1, t2, t3 = (3 - q) ** 5, -6 * (2 - q) ** 5, 15 * (1 - q) ** 5
        f = np.where(q < 1, t1 + t2 + t3, np.where(q < 2, t1 + t2, np.where(q < 3, t1, 0.0))) / (120 * math.pi)
    want = sc["uint"] + 3.0 * f / hk**3
    assert np.allclose(m.get(0, "uint"), want, rtol=1e-13, atol=0)
    m.set_value_in_sphere("uint", 7.0, c, 3 * sc["dr"])
    assert np.array_equal(m.get(0, "uint") == 7.0, r * r < (3 * sc["dr"]) ** 2) or np.array_equal(
        m.get(0, "uint") == 7.0, ((x - np.array(c)) ** 2).sum(1) < (3 * sc["dr"]) ** 2)


def test_disc_mc_is_layout_independent_and_sane():
    """the Monte-Carlo disc: the same objects whatever the patch grid (counter-based draws), radii inside
    [r_in, r_out], Keplerian speeds, surface density ~ r^-p"""
    sc = S.disc(1000, "M4")
    args = dict(npart=200000, seed=1234, r_in=1.0, r_out=3.0, p=1.0, q=0.25, H_r_in=0.05, disc_mass=0.01)
    out = []
    for grid in [(1, 1, 1), (2, 2, 1)]:
        m, _ = empty_model(sc, grid)
        m.set_particle_mass(args["disc_mass"] / args["npart"])
        n = m.add_disc_mc(**args)
        assert n == args["npart"]
        x = np.concatenate([m.get(ip, "xyz") for ip in range(m.patch_count) if m.patch_size(ip)])
        v = np.concatenate([m.get(ip, "vxyz") for ip in range(m.patch_count) if m.patch_size(ip)])
        h = np.concatenate([m.get(ip, "hpart") for ip in range(m.patch_count) if m.patch_size(ip)])
        order = np.lexsort((x[:, 2], x[:, 1], x[:, 0]))
        out.append((x[order], v[order], h[order]))
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    x, v, h = out[0]
    r = np.hypot(x[:, 0], x[:, 1])
    assert r.min() >= 1.0 and r.max() <= 3.0 and (h > 0).all() and np.isfinite(h).all()
    assert np.allclose(np.hypot(v[:, 0], v[:, 1]), 1 / np.sqrt(r), rtol=1e-12)
    # Sigma ~ 1 / r: the number of objects per unit radius is flat
    cnt, _ = np.histogram(r, bins=8, range=(1.0, 3.0))
    assert cnt.std() / cnt.mean() < 0.03
    # vertical scale: <z^2> = H^2 with H = 0.05 r^(1.25)
    H = 0.05 * r**1.25
    assert abs((x[:, 2] ** 2 / H**2).mean() - 1) < 0.02
