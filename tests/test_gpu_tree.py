"""GPU parity, stage level: the C-ABI tree / neighbour-cache / h-iteration entry points of
libshamb200.so against the CPU oracle on the same seeded inputs (bit-exact), against the reference's
golden vectors, and size-independent properties at bench sizes.

Device memory is carried by torch tensors (plumbing only); every compute call goes through
include/shamb200.h."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import pyoracle as po  # noqa: E402  (checker only)
from shamrock_b200 import _capi  # noqa: E402

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_goldens.json")))


@pytest.fixture(scope="module")
def ctx():
    c = _capi.Context(0)
    yield c
    c.close()


def dev(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    torch.cuda.synchronize()
    return t


def fetch(ptr, n, dtype):
    """copy n elements from a raw device pointer (ctx arena) to numpy"""
    import ctypes as C

    out = np.empty(n, dtype=dtype)
    if n:
        cudart = C.CDLL("libcudart.so")
        rc = cudart.cudaMemcpy(out.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), C.c_size_t(out.nbytes), 2)
        assert rc == 0
    return out


def tree_arrays(tv):
    L, I, P2 = tv.leaf_count, tv.int_count, tv.morton_count
    return dict(
        sorted_morton=fetch(tv.d_sorted_morton, P2, np.uint32),
        sort_index_map=fetch(tv.d_sort_index_map, P2, np.uint32),
        reduc_index_map=fetch(tv.d_reduc_index_map, L + 2, np.uint32),
        reduced_morton=fetch(tv.d_reduced_morton, L, np.uint32),
        lchild_id=fetch(tv.d_lchild_id, I, np.uint32),
        rchild_id=fetch(tv.d_rchild_id, I, np.uint32),
        endrange=fetch(tv.d_endrange, I, np.uint32),
        lchild_flag=fetch(tv.d_lchild_flag, I, np.uint8),
        rchild_flag=fetch(tv.d_rchild_flag, I, np.uint8),
        aabb_min=fetch(tv.d_aabb_min, (I + L) * 3, np.float64).reshape(-1, 3),
        aabb_max=fetch(tv.d_aabb_max, (I + L) * 3, np.float64).reshape(-1, 3),
    )


def positions(kind, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.uniform(0, 1, (n, 3))
    if kind == "clustered":  # r -> r |r|^2 : heavy Morton ties near the centre
        r = rng.uniform(-1, 1, (n, 3))
        return 0.5 + 0.5 * r * (np.linalg.norm(r, axis=1, keepdims=True) ** 2)
    if kind == "identical":
        return np.full((n, 3), 0.25)
    if kind == "lattice":
        m = int(round(n ** (1 / 3))) + 1
        g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n]
        return (g + 0.5) / m
    raise ValueError(kind)


CASES = [("uniform", 1), ("uniform", 2), ("uniform", 3), ("uniform", 17), ("uniform", 1000),
         ("uniform", 4096), ("uniform", 4097), ("clustered", 30000), ("identical", 100), ("lattice", 50000),
         ("uniform", 300000)]


@pytest.mark.parametrize("kind,n", CASES)
@pytest.mark.parametrize("level", [0, 3])
def test_tree_build_bitonic_bit_exact(ctx, kind, n, level):
    xyz = positions(kind, n, 1234 + n)
    bb = ([0.0, 0.0, 0.0], [1.0, 1.0, 1.0])
    ref = po.Tree(xyz, *bb, level, bits=32)
    tv = ctx.tree_build(dev(xyz), n, *bb, reduction_level=level, sort_mode="bitonic")
    ctx.synchronize()
    assert (tv.obj_cnt, tv.morton_count, tv.leaf_count, tv.int_count) == (
        ref.obj_cnt, ref.morton_count, ref.leaf_count, ref.int_count)
    got = tree_arrays(tv)
    for k, v in got.items():
        r = ref.get(k)
        assert v.shape == r.shape, k
        assert np.array_equal(v, r), f"{k}: first mismatch at {np.argwhere(v != r)[:3].tolist()}"


@pytest.mark.parametrize("kind,n", CASES)
def test_tree_build_radix_same_tree_modulo_ties(ctx, kind, n):
    """stable radix sort: same sorted keys, same tree topology and AABBs (F3: the topology is built
    from the unique reduced codes); inside an equal-key run the objects are in input order."""
    xyz = positions(kind, n, 99 + n)
    bb = ([0.0, 0.0, 0.0], [1.0, 1.0, 1.0])
    ref = po.Tree(xyz, *bb, 3, bits=32)
    tv = ctx.tree_build(dev(xyz), n, *bb, reduction_level=3, sort_mode="radix")
    ctx.synchronize()
    got = tree_arrays(tv)
    for k in ("sorted_morton", "reduc_index_map", "reduced_morton", "lchild_id", "rchild_id", "endrange",
              "lchild_flag", "rchild_flag", "aabb_min", "aabb_max"):
        assert np.array_equal(got[k], ref.get(k)), k
    codes = po.morton_codes(xyz, *bb, tv.morton_count, bits=32)
    order = np.argsort(codes[:n], kind="stable")
    assert np.array_equal(got["sort_index_map"][:n], order.astype(np.uint32))


def test_tree_golden_u32_shift(ctx):
    """the reference's 14-position golden (MortonCodeSetTests.cpp:25-77) holds u64 codes; the u32 codes
    of the same positions are the top 30 of the 63 bits.  Sorted order of distinct codes must agree."""
    f = "src/tests/shamtree/MortonCodeSetTests.cpp"
    pos = np.array(G[f]["<file>"]["partpos"]["value"])
    tv = ctx.tree_build(dev(pos), len(pos), [0, 0, 0], [1, 1, 1], reduction_level=0, sort_mode="bitonic")
    ctx.synchronize()
    got = tree_arrays(tv)
    c64 = np.array(G[f]["<file>"]["test_mortons"]["value"], dtype=np.uint64)[: len(pos)]
    c32 = (c64 >> np.uint64(33)).astype(np.uint32)
    assert np.array_equal(np.sort(c32), got["sorted_morton"][: len(pos)])
    assert (got["sorted_morton"][len(pos):] == 0xFFFFFFFF).all()


def test_tree_auto_bbox(ctx):
    xyz = positions("uniform", 5000, 5) * 3 - 1
    tv = ctx.tree_build(dev(xyz), len(xyz))
    ctx.synchronize()
    lo, hi = xyz.min(0), xyz.max(0)
    assert list(tv.bmin) == list(np.nextafter(lo, -np.inf))
    assert list(tv.bmax) == list(np.nextafter(hi, np.inf))
    ref = po.Tree(xyz, list(tv.bmin), list(tv.bmax), 3, bits=32)
    got = tree_arrays(tv)
    for k, v in got.items():
        assert np.array_equal(v, ref.get(k)), k


def test_tree_errors(ctx):
    with pytest.raises(_capi.ShamB200Error, match="obj_cnt is 0"):
        ctx.tree_build(dev(np.zeros((1, 3))), 0, [0, 0, 0], [1, 1, 1])


@pytest.mark.parametrize("kernel,two_stage", [("M4", True), ("M4", False), ("M6", True), ("M6", False)])
@pytest.mark.parametrize("kind,n", [("uniform", 20000), ("clustered", 8000), ("lattice", 15000), ("uniform", 40)])
def test_neighbour_cache_bit_exact(ctx, kernel, two_stage, kind, n):
    R = {"M4": 2.0, "M6": 3.0}[kernel]
    xyz = positions(kind, n, 77 + n)
    rng = np.random.default_rng(n)
    h = (0.7 / n ** (1 / 3)) * rng.uniform(0.8, 1.3, n)
    n_real = n - n // 5  # the last fifth plays the ghosts: objects without a list
    bb = ([0.0, 0.0, 0.0], [1.0, 1.0, 1.0])
    ref = po.Tree(xyz, *bb, 3, bits=32)
    rint_ref = ref.field_max(h, 1.1)
    cref = ref.neigh_cache(h, n_real, R, 1.1, two_stage)
    dx, dh = dev(xyz), dev(h)
    tv = ctx.tree_build(dx, n, *bb, reduction_level=3, sort_mode="bitonic")
    rint = torch.empty(tv.leaf_count + tv.int_count, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ctx.tree_field_max(tv, dh, 1.1, rint)
    cv = ctx.neigh_cache_build(tv, dx, dh, rint, n_real, R, 1.1, two_stage)
    ctx.synchronize()
    assert np.array_equal(rint.cpu().numpy(), rint_ref)
    assert cv.obj_cnt == n_real and cv.sum_neigh_cnt == len(cref["index_neigh_map"])
    assert np.array_equal(fetch(cv.d_cnt_neigh, n_real, np.uint32), cref["cnt_neigh"])
    assert np.array_equal(fetch(cv.d_scanned_cnt, n_real, np.uint32), cref["scanned_cnt"])
    assert np.array_equal(fetch(cv.d_index_neigh_map, cv.sum_neigh_cnt, np.uint32), cref["index_neigh_map"])


def _cache_case(c, xyz, h, n_real, kernel, sort_mode="bitonic"):
    R = {"M4": 2.0, "M6": 3.0}[kernel]
    n = len(xyz)
    bb = ([0.0, 0.0, 0.0], [1.0, 1.0, 1.0])
    ref = po.Tree(xyz, *bb, 3, bits=32)
    ref.field_max(h, 1.1)
    cref = ref.neigh_cache(h, n_real, R, 1.1, True)
    dx, dh = dev(xyz), dev(h)
    tv = c.tree_build(dx, n, *bb, reduction_level=3, sort_mode=sort_mode)
    got = tree_arrays(tv)
    for k, v in got.items():
        assert np.array_equal(v, ref.get(k)), k
    rint = torch.empty(tv.leaf_count + tv.int_count, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    c.tree_field_max(tv, dh, 1.1, rint)
    cv = c.neigh_cache_build(tv, dx, dh, rint, n_real, R, 1.1, True)
    c.synchronize()
    assert cv.sum_neigh_cnt == len(cref["index_neigh_map"])
    assert np.array_equal(fetch(cv.d_cnt_neigh, n_real, np.uint32), cref["cnt_neigh"])
    assert np.array_equal(fetch(cv.d_index_neigh_map, cv.sum_neigh_cnt, np.uint32), cref["index_neigh_map"])
    return c.neigh_cache_stats()


@pytest.mark.parametrize("kind,n", [("uniform", 1 << 21), ("lattice", 3 << 20)])
def test_tree_and_cache_millions_bit_exact(ctx, kind, n):
    """the sizes the benchmark runs at, against the oracle: tree (codes, permutation, topology, boxes) and the
    ordered neighbour lists of 2 - 3 Mi objects, ~77 neighbours each, bit for bit (radix tiles over hundreds of
    CTAs, 10^5 leaf groups, 1.6 - 2.4 x 10^8 list entries)"""
    xyz = positions(kind, n, 4242)
    rng = np.random.default_rng(n)
    if kind == "lattice":  # jittered lattice: the benchmark's kind of input
        xyz = np.clip(xyz + rng.uniform(-0.1, 0.1, xyz.shape) / n ** (1 / 3), 0.0, 1.0 - 1e-12)
    h = (1.2 / n ** (1 / 3)) * rng.uniform(0.9, 1.1, n)
    st = _cache_case(ctx, xyz, h, n, "M4")
    assert st["K"] > 60 * n


def test_neighbour_cache_candidate_array_regrow():
    """a fresh context sizes the candidate-entry array for ~24 candidate leaves per leaf; M6 with a large h needs
    more: the search reads the exact need back, regrows and repeats (attempts > 1) — lists still bit-exact"""
    c = _capi.Context(0)
    n = 60000
    xyz = positions("uniform", n, 11)
    h = np.full(n, 1.6 / n ** (1 / 3))
    st = _cache_case(c, xyz, h, n, "M6")
    assert st["attempts"] > 1, st
    st2 = _cache_case(c, xyz, h, n, "M6")  # the second search of the same context fits at once
    assert st2["attempts"] == 1, st2
    c.close()


def test_neighbour_cache_global_frontier_groups_at_scale():
    """half a million objects, three of them with an h of a tenth of the box: thousands of leaf groups overflow
    the shared-memory frontier and are walked with a frontier in global memory (group_walk_kernel<true>)"""
    n = 500000
    xyz = positions("uniform", n, 5)
    rng = np.random.default_rng(9)
    h = (1.0 / n ** (1 / 3)) * rng.uniform(0.9, 1.1, n)
    h[[17, 4711, 200000]] = [0.1, 0.12, 0.08]
    c = _capi.Context(0)
    st = _cache_case(c, xyz, h, n, "M4")
    assert st["over_groups"] > 0, st
    c.close()


@pytest.mark.parametrize("kernel", ["M4", "M6"])
def test_neighbour_cache_outlier_h(ctx, kernel):
    """A few particles with a smoothing length of the size of the box (what an unconverged h iteration leaves
    behind): their leaf groups see thousands of candidate leaves, far beyond the shared-memory frontier of the
    tree walk — those groups are walked with a frontier in global memory and the lists stay bit-exact."""
    R = {"M4": 2.0, "M6": 3.0}[kernel]
    n = 30000
    xyz = positions("uniform", n, 5)
    rng = np.random.default_rng(9)
    h = (0.7 / n ** (1 / 3)) * rng.uniform(0.8, 1.3, n)
    h[[17, 4711, 20000]] = [0.25, 0.4, 0.1]
    bb = ([0.0, 0.0, 0.0], [1.0, 1.0, 1.0])
    ref = po.Tree(xyz, *bb, 3, bits=32)
    ref.field_max(h, 1.1)
    cref = ref.neigh_cache(h, n, R, 1.1, True)
    dx, dh = dev(xyz), dev(h)
    tv = ctx.tree_build(dx, n, *bb, reduction_level=3, sort_mode="bitonic")
    rint = torch.empty(tv.leaf_count + tv.int_count, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ctx.tree_field_max(tv, dh, 1.1, rint)
    cv = ctx.neigh_cache_build(tv, dx, dh, rint, n, R, 1.1, True)
    ctx.synchronize()
    assert cref["cnt_neigh"].max() > 5000  # the outliers really see a large part of the box
    assert cv.sum_neigh_cnt == len(cref["index_neigh_map"])
    assert np.array_equal(fetch(cv.d_cnt_neigh, n, np.uint32), cref["cnt_neigh"])
    assert np.array_equal(fetch(cv.d_index_neigh_map, cv.sum_neigh_cnt, np.uint32), cref["index_neigh_map"])


@pytest.mark.parametrize("n_same", [40, 1500])
def test_neighbour_cache_coincident_particles(ctx, n_same):
    """Many particles at exactly the same position: one Morton code, one huge leaf.  Exercises the batches
    of 32 particles per leaf, candidate windows larger than the shared-memory rank cache (entries of more
    than 1024 ranks are streamed straight from the entry) and the re-test fill when the ballots do not fit."""
    n, R = 3000, 2.0
    xyz = positions("uniform", n, 11)
    xyz[100:100 + n_same] = xyz[100]
    xyz[2000:2000 + n_same // 2] = xyz[2000] + 1e-9  # a second clump, distinct positions inside one cell
    h = np.full(n, 0.05)
    h[100:100 + n_same:7] = 0.09
    bb = ([0.0, 0.0, 0.0], [1.0, 1.0, 1.0])
    ref = po.Tree(xyz, *bb, 3, bits=32)
    ref.field_max(h, 1.1)
    cref = ref.neigh_cache(h, n, R, 1.1, True)
    dx, dh = dev(xyz), dev(h)
    tv = ctx.tree_build(dx, n, *bb, reduction_level=3, sort_mode="bitonic")
    rint = torch.empty(tv.leaf_count + tv.int_count, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ctx.tree_field_max(tv, dh, 1.1, rint)
    cv = ctx.neigh_cache_build(tv, dx, dh, rint, n, R, 1.1, True)
    ctx.synchronize()
    assert cref["cnt_neigh"].max() >= n_same
    assert cv.sum_neigh_cnt == len(cref["index_neigh_map"])
    assert np.array_equal(fetch(cv.d_cnt_neigh, n, np.uint32), cref["cnt_neigh"])
    assert np.array_equal(fetch(cv.d_index_neigh_map, cv.sum_neigh_cnt, np.uint32), cref["index_neigh_map"])


def test_neighbour_cache_brute_force(ctx):
    """independent of the oracle: list == all pairs passing the accept test, in sorted-Morton rank"""
    n, R, tol = 3000, 2.0, 1.1
    xyz = positions("uniform", n, 3)
    h = np.full(n, 0.06) * np.random.default_rng(1).uniform(0.9, 1.2, n)
    dx, dh = dev(xyz), dev(h)
    tv = ctx.tree_build(dx, n, [0, 0, 0], [1, 1, 1], reduction_level=3, sort_mode="bitonic")
    rint = torch.empty(tv.leaf_count + tv.int_count, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ctx.tree_field_max(tv, dh, tol, rint)
    cv = ctx.neigh_cache_build(tv, dx, dh, rint, n, R, tol, True)
    ctx.synchronize()
    order = fetch(tv.d_sort_index_map, n, np.uint32)
    rank = np.empty(n, dtype=np.int64)
    rank[order] = np.arange(n)
    d = xyz[:, None, :] - xyz[None, :, :]
    r2 = d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]
    ra = h * tol
    lim = ra * ra * (R * R)
    acc = ~((r2 > lim[:, None]) & (r2 > lim[None, :]))
    cnt = fetch(cv.d_cnt_neigh, n, np.uint32)
    sc = fetch(cv.d_scanned_cnt, n, np.uint32)
    lst = fetch(cv.d_index_neigh_map, cv.sum_neigh_cnt, np.uint32)
    assert np.array_equal(cnt, acc.sum(1).astype(np.uint32))
    for a in range(0, n, 37):
        nb = np.nonzero(acc[a])[0]
        nb = nb[np.argsort(rank[nb])]
        assert np.array_equal(lst[sc[a]: sc[a] + cnt[a]], nb.astype(np.uint32))


def test_h_iteration_reference_golden(ctx):
    """IterateSmoothingLengthDensityTests.cpp:250-366 through the C ABI: 4x4x4 unit lattice, all-pairs
    cache, M4, m = 1, tolerances 1.2 / 1.2 — final h, eps and the per-sweep min/max sequences."""
    f = "src/tests/shammodels/sph/modules/IterateSmoothingLengthDensityTests.cpp"
    t = "shammodels/sph/modules/IterateSmoothingLengthDensity"
    gv = lambda k: np.array(G[f][t][k]["value"])
    pos = np.array([[i, j, k] for i in range(4) for j in range(4) for k in range(4)], dtype=np.float64)
    n = len(pos)
    cv = _capi.CsrView()
    cnt, sc = dev(np.full(n, n, dtype=np.uint32)), dev((np.arange(n) * n).astype(np.uint32))
    idx = dev(np.tile(np.arange(n, dtype=np.uint32), n))
    cv.obj_cnt, cv.sum_neigh_cnt = n, n * n
    cv.d_cnt_neigh, cv.d_scanned_cnt, cv.d_index_neigh_map = cnt.data_ptr(), sc.data_ptr(), idx.data_ptr()
    dx = dev(pos)
    h_new = dev(np.full(n, 0.1))
    eps = dev(np.zeros(n))
    seq = dict(eps_min=[], eps_max=[], h_min=[], h_max=[])
    done = False
    for outer in range(50):
        h_old = h_new.clone()
        eps.fill_(10000000.0)
        torch.cuda.synchronize()
        max_eps = 1e7
        for inner in range(10):
            ctx.h_iterate("M4", cv, dx, h_old, h_new, eps, 1.0, 1.2, 1.2)
            ctx.synchronize()
            e, hh = eps.cpu().numpy(), h_new.cpu().numpy()
            seq["eps_min"].append(e.min()), seq["eps_max"].append(e.max())
            seq["h_min"].append(hh.min()), seq["h_max"].append(hh.max())
            max_eps = e.max()
            if max_eps < 1e-6:
                break
        if eps.min().item() == -1:
            continue
        if max_eps < 1e-6:
            done = True
            break
    assert done
    tol = 1e-6  # tolerance of the reference test
    assert np.abs(h_new.cpu().numpy() - gv("expected_h_vec_end")).max() <= tol
    assert np.abs(eps.cpu().numpy() - gv("expected_eps_vec_end")).max() <= tol
    for k in seq:
        exp = gv("expected_sequence_" + k)
        assert len(seq[k]) == len(exp) and np.abs(np.array(seq[k]) - exp).max() <= tol, k


@pytest.mark.parametrize("kernel", ["M4", "M6"])
def test_h_iterate_and_omega_vs_oracle(ctx, kernel):
    from tests import scenarios as S

    n = 6000
    R = S.RKERN[kernel]
    xyz = positions("uniform", n, 11)
    pm = 1.0 / n
    h0 = np.full(n, S.HFACT[kernel] * (1.0 / n) ** (1 / 3)) * np.random.default_rng(5).uniform(0.9, 1.1, n)
    bb = ([0.0, 0.0, 0.0], [1.0, 1.0, 1.0])
    ref = po.Tree(xyz, *bb, 3, bits=32)
    ref.field_max(h0, 1.1)
    cref = ref.neigh_cache(h0, n, R, 1.1, True)
    dx, dh0 = dev(xyz), dev(h0)
    tv = ctx.tree_build(dx, n, *bb)
    rint = torch.empty(tv.leaf_count + tv.int_count, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ctx.tree_field_max(tv, dh0, 1.1, rint)
    cv = ctx.neigh_cache_build(tv, dx, dh0, rint, n, R, 1.1, True)
    h_ref, e_ref = h0.copy(), np.full(n, 100.0)
    h_new, eps = dev(h0), dev(np.full(n, 100.0))
    for sweep in range(4):
        po.h_iterate(kernel, cref, xyz, h0, h_ref, e_ref, pm, 1.1, 1.1)
        ctx.h_iterate(kernel, cv, dx, dh0, h_new, eps, pm, 1.1, 1.1)
        ctx.synchronize()
        assert np.array_equal(h_new.cpu().numpy(), h_ref), sweep
        assert np.array_equal(eps.cpu().numpy(), e_ref), sweep
    # the loop entry point does the same sweeps + the convergence reduction
    h2, e2 = dev(h0), dev(np.full(n, 100.0))
    torch.cuda.synchronize()
    r = ctx.h_iterate_loop(kernel, cv, dx, dh0, h2, e2, pm, 1.1, 1.1, 1e-6, 50)
    e2n = e2.cpu().numpy()
    assert r["max_eps"] == e2n.max() and r["min_eps"] == e2n.min()
    assert r["max_eps"] < 1e-6 or r["sweeps"] == 50


@pytest.mark.parametrize("n", [1 << 20, (1 << 22) + 12345])
@pytest.mark.parametrize("mode", ["bitonic", "radix"])
def test_tree_large_properties(ctx, n, mode):
    """bench-size inputs (config C2): size-independent properties instead of the oracle."""
    g = torch.Generator(device="cuda").manual_seed(n)
    xyz = torch.rand((n, 3), dtype=torch.float64, device="cuda", generator=g)
    torch.cuda.synchronize()
    tv = ctx.tree_build(xyz, n, [0, 0, 0], [1, 1, 1], reduction_level=3, sort_mode=mode)
    ctx.synchronize()
    a = tree_arrays(tv)
    L, I = tv.leaf_count, tv.int_count
    sm = a["sorted_morton"]
    assert (np.diff(sm.astype(np.int64)) >= 0).all()  # sortedness
    assert (sm[n:] == 0xFFFFFFFF).all()
    perm = a["sort_index_map"][:n]
    assert np.array_equal(np.sort(perm), np.arange(n, dtype=np.uint32))  # a permutation
    # codes of the permuted positions are the sorted codes (sort carries the values with the keys)
    codes = po.morton_codes(xyz.cpu().numpy(), [0, 0, 0], [1, 1, 1], tv.morton_count, bits=32)
    assert np.array_equal(codes[perm], sm[:n])
    rim = a["reduc_index_map"]
    assert rim[L] == n and rim[L + 1] == 0 and rim[0] == 0 and (np.diff(rim[: L + 1].astype(np.int64)) > 0).all()
    assert (np.diff(a["reduced_morton"].astype(np.int64)) > 0).all()  # unique leaf codes
    assert np.array_equal(a["reduced_morton"], sm[rim[:L]])
    # every node except the root has exactly one parent; leaves and internal cells all covered
    lc = a["lchild_id"].astype(np.int64) + I * a["lchild_flag"].astype(np.int64)
    rc = a["rchild_id"].astype(np.int64) + I * a["rchild_flag"].astype(np.int64)
    tgt = np.concatenate([lc, rc])
    assert np.array_equal(np.sort(tgt), np.arange(1, I + L))
    # root box = box of all points; every child box inside its parent's
    pos = xyz.cpu().numpy()
    assert np.array_equal(a["aabb_min"][0], pos.min(0)) and np.array_equal(a["aabb_max"][0], pos.max(0))
    assert np.array_equal(a["aabb_min"][:I], np.minimum(a["aabb_min"][lc], a["aabb_min"][rc]))
    assert np.array_equal(a["aabb_max"][:I], np.maximum(a["aabb_max"][lc], a["aabb_max"][rc]))


def test_pyshamrock_tree_surface():
    """shamrock.tree.CLBVH_*().rebuild_from_positions / get_*_cell_count (shampylib/src/pyShamtree.cpp:28-60)
    through shamrock_b200.pyshamrock, against the oracle"""
    from shamrock_b200 import pyshamrock as shamrock

    n = 20000
    xyz = positions("uniform", n, 3)
    buf = shamrock.backends.DeviceBuffer_f64_3()
    buf.resize(n)
    buf.copy_from_stdvec(xyz)
    assert buf.get_size() == n
    bvh = shamrock.tree.CLBVH_u32_f64_3()
    with pytest.raises(RuntimeError):
        bvh.get_leaf_cell_count()
    bvh.rebuild_from_positions(buf, shamrock.math.AABB_f64_3((0.0, 0.0, 0.0), (1.0, 1.0, 1.0)), 3)
    ref = po.Tree(xyz, [0, 0, 0], [1, 1, 1], 3, bits=32)
    assert bvh.get_leaf_cell_count() == ref.leaf_count
    assert bvh.get_internal_cell_count() == ref.int_count
    assert bvh.get_total_cell_count() == ref.leaf_count + ref.int_count
