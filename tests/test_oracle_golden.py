"""Pin the CPU oracle against every golden vector the reference's own unit tests hold for the hot
path.  Vectors: tests/golden/reference_goldens.json (extracted from
/root/reference/src/tests by tests/golden/extract_goldens.py)."""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as po

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_goldens.json")))
BB = ([0.0, 0.0, 0.0], [1.0, 1.0, 1.0])


def g(file, test, var):
    return G[f"src/tests/{file}"][test][var]["value"]


def u64(a):
    return np.array(a, dtype=np.uint64)


# ---- MortonCodeSetTests.cpp:22-125 ----------------------------------------------------------
def test_morton_code_set_and_sorted_set():
    f, t = "shamtree/MortonCodeSetTests.cpp", "<file>"
    pos = np.array(g(f, t, "partpos"))
    codes = po.morton_codes(pos, *BB, 16, bits=64)
    assert (codes == u64(g(f, t, "test_mortons"))).all()
    k, v = po.sort_by_key(codes, bits=64)
    assert (k == u64(g(f, t, "test_mortons_sorted"))).all()
    assert v.tolist() == g(f, t, "index_map_obj_idx")  # pins the tie order of the bitonic network


# ---- MortonReducedSetTests.cpp:26-248 -------------------------------------------------------
@pytest.mark.parametrize("t", ["shamtree/MortonReducedSet", "shamtree/MortonReducedSet(single cell)"])
def test_morton_reduced_set(t):
    f = "shamtree/MortonReducedSetTests.cpp"
    pos = np.array(g(f, t, "partpos"))
    codes = po.morton_codes(pos, *BB, 16, bits=64)
    assert (codes == u64(g(f, t, "test_mortons"))).all()
    k, v = po.sort_by_key(codes, bits=64)
    assert (k == u64(g(f, t, "test_mortons_sorted"))).all()
    assert v.tolist() == g(f, t, "index_map_obj_idx")
    imap, lc = po.reduction(k, len(pos), 2, bits=64)
    assert imap.tolist() == g(f, t, "buf_reduc_index_map")
    assert (k[imap[:lc]] == u64(g(f, t, "reduced_morton_codes"))).all()


# ---- KarrasRadixTreeTests.cpp:19-107 --------------------------------------------------------
@pytest.mark.parametrize("t", ["shamtree/KarrasRadixTree", "shamtree/KarrasRadixTree(one-cell)"])
def test_karras_radix_tree(t):
    f = "shamtree/KarrasRadixTreeTests.cpp"
    r = po.karras(u64(g(f, t, "test_morton_codes")), bits=64)
    for k in ("lchild_id", "rchild_id", "lchild_flag", "rchild_flag", "endrange"):
        assert r[k].tolist() == g(f, t, "expected_" + k), k


def test_karras_radix_tree_u32_same_topology():
    """SPH uses u32 codes (SolverConfig.hpp:443); the same 12 codes shifted to 30 bits must give
    the same topology (the algorithm only looks at common-prefix lengths)."""
    f, t = "shamtree/KarrasRadixTreeTests.cpp", "shamtree/KarrasRadixTree"
    c64 = u64(g(f, t, "test_morton_codes"))
    c32 = (c64 >> np.uint64(33)).astype(np.uint32)
    assert len(set(c32.tolist())) == len(c32)
    r = po.karras(c32, bits=32)
    for k in ("lchild_id", "rchild_id", "lchild_flag", "rchild_flag", "endrange"):
        assert r[k].tolist() == g(f, t, "expected_" + k), k


# ---- KarrasRadixTreeAABBTests.cpp:24-396 ----------------------------------------------------
@pytest.mark.parametrize(
    "t,level", [("shamtree/KarrasRadixTreeAABB", 1), ("shamtree/KarrasRadixTreeAABB(one-cell)", 5)]
)
def test_karras_radix_tree_aabb(t, level):
    f = "shamtree/KarrasRadixTreeAABBTests.cpp"
    pos = np.array(g(f, t, "partpos"))
    tr = po.Tree(pos, *BB, level, bits=64, morton_count=16)  # the test pads to 16 explicitly
    assert (tr.get("sorted_morton") == u64(g(f, t, "test_mortons_sorted"))).all()
    assert tr.get("sort_index_map").tolist() == g(f, t, "index_map_obj_idx")
    assert tr.get("reduc_index_map").tolist() == g(f, t, "buf_reduc_index_map")
    assert (tr.get("reduced_morton") == u64(g(f, t, "reduced_morton_codes"))).all()
    for k in ("lchild_id", "rchild_id", "lchild_flag", "rchild_flag", "endrange"):
        assert tr.get(k).tolist() == g(f, t, "expected_" + k), k
    assert (tr.get("aabb_min") == np.array(g(f, t, "aabb_min")).reshape(-1, 3)).all()
    assert (tr.get("aabb_max") == np.array(g(f, t, "aabb_max")).reshape(-1, 3)).all()


# ---- KarrasRadixTreeFieldTests.cpp:24-177 ---------------------------------------------------
def test_karras_radix_tree_field():
    f, t = "shamtree/KarrasRadixTreeFieldTests.cpp", "shamtree/KarrasRadixTreeField"
    pos = np.array(g(f, t, "partpos"))
    tr = po.Tree(pos, *BB, 1, bits=64)
    assert tr.get("sort_index_map").tolist() == g(f, t, "index_map_obj_idx")
    assert tr.get("reduc_index_map").tolist() == g(f, t, "buf_reduc_index_map")
    for k in ("lchild_id", "rchild_id", "lchild_flag", "rchild_flag", "endrange"):
        assert tr.get(k).tolist() == g(f, t, "expected_" + k), k
    res = tr.field_max(np.array(g(f, t, "field_values")), 1.0)
    assert res.tolist() == g(f, t, "expected_result")


# ---- CLBVHObjectIteratorTests.cpp:25-439 ----------------------------------------------------
@pytest.mark.parametrize(
    "t,level,nint", [("shamtree/LCBVHObjectIterator", 1, 6), ("shamtree/LCBVHObjectIterator(one-cell)", 8, 0)]
)
def test_clbvh_object_iterator(t, level, nint):
    f = "shamtree/CLBVHObjectIteratorTests.cpp"
    pos = np.array(g(f, t, "partpos"))
    tr = po.Tree(pos, *BB, level, bits=64)
    assert tr.int_count == nint
    c = tr.box_query(-1.0)  # "find everything"
    assert c["cnt_neigh"].tolist() == g(f, t, "expected_counts")
    assert c["index_neigh_map"].tolist() == g(f, t, "expected_neigh")
    c = tr.box_query(0.15)  # "find within a box around particles"
    assert c["cnt_neigh"].tolist() == g(f, t, "expected_counts#2")
    assert c["index_neigh_map"].tolist() == g(f, t, "expected_neigh#2")


# ---- IterateSmoothingLengthDensityTests.cpp:250-366 -------------------------------------------
def test_iterate_smoothing_length_density():
    f = "shammodels/sph/modules/IterateSmoothingLengthDensityTests.cpp"
    t = "shammodels/sph/modules/IterateSmoothingLengthDensity"
    pos = np.array([[i, j, k] for i in range(4) for j in range(4) for k in range(4)], dtype=np.float64)
    n = len(pos)
    cache = dict(
        cnt_neigh=np.full(n, n, dtype=np.uint32),
        scanned_cnt=(np.arange(n) * n).astype(np.uint32),
        index_neigh_map=np.tile(np.arange(n, dtype=np.uint32), n),
    )
    h_new = np.full(n, 0.1)
    eps = np.zeros(n)
    seq = dict(eps_min=[], eps_max=[], h_min=[], h_max=[])
    done = False
    for outer in range(50):  # driver loop of the reference test, lines 143-215
        h_old = h_new.copy()
        eps[:] = 10000000.0
        max_eps = 1e7
        for inner in range(10):
            po.h_iterate("M4", cache, pos, h_old, h_new, eps, 1.0, 1.2, 1.2)
            seq["eps_min"].append(eps.min())
            seq["eps_max"].append(eps.max())
            seq["h_min"].append(h_new.min())
            seq["h_max"].append(h_new.max())
            assert ((eps >= 0) | (eps == -1.0)).all()
            max_eps = eps.max()
            if max_eps < 1e-6:
                break
        if eps.min() == -1:
            continue
        if max_eps < 1e-6:
            done = True
            break
    assert done
    tol = 1e-6  # the tolerance of the reference test (line 224)
    assert np.abs(h_new - np.array(g(f, t, "expected_h_vec_end"))).max() <= tol
    assert np.abs(eps - np.array(g(f, t, "expected_eps_vec_end"))).max() <= tol
    for k in ("eps_min", "eps_max", "h_min", "h_max"):
        exp = np.array(g(f, t, "expected_sequence_" + k))
        got = np.array(seq[k])
        assert got.shape == exp.shape, k
        assert np.abs(got - exp).max() <= tol, k
    # the oracle is IEEE / no-FMA like the CPU run that produced the goldens: much tighter in fact
    assert np.abs(h_new - np.array(g(f, t, "expected_h_vec_end"))).max() <= 1e-14


# ---- sphkernelsTests.cpp:25-160 (identities; no stored vectors) --------------------------------
@pytest.mark.parametrize("kern,R", [("M4", 2.0), ("M6", 3.0)])
def test_sph_kernel_identities(kern, R):
    q = np.linspace(0, R * 1.2, 2401)
    f = np.array([po.kernel_eval(kern, "f", x) for x in q])
    df = np.array([po.kernel_eval(kern, "df", x) for x in q])
    assert (f[q >= R] == 0).all() and (df[q >= R] == 0).all()  # compact support
    norm = {"M4": 1 / np.pi, "M6": 1 / (120 * np.pi)}[kern]
    integ = np.trapezoid(4 * np.pi * q**2 * f * norm, q)
    assert abs(integ - 1) < 1e-5  # ∫ W d^3x = 1
    num = np.gradient(f, q)
    assert np.abs(num[2:-2] - df[2:-2]).max() < 2e-2 * max(1, np.abs(df).max())  # df = f'
    h = 0.7
    for r in (0.1, 0.5, 1.3):  # W(r,h) = norm f(r/h)/h^3, dW = norm df(r/h)/h^4
        assert po.kernel_eval(kern, "W_3d", r, h) == norm * po.kernel_eval(kern, "f", r / h) / (h * h * h)
        assert po.kernel_eval(kern, "dW_3d", r, h) == norm * po.kernel_eval(kern, "df", r / h) / (h * h * h * h)


def test_bit_interleave_against_the_reference_code_itself():
    """shammath/sfc/bmi.hpp compiled where it lies (oracle/_ref/bmi_ref): the Morton codes of the oracle are
    expand(x) * 4 + expand(y) * 2 + expand(z) of the reference's own expand_bits (morton.hpp:59-64,127-132), and the
    bit spreading of the Hilbert restatement (oracle/load_balance.py) is the reference's expand_bits<u64, 2>"""
    from oracle import io_formats as io
    from oracle import load_balance as lb
    from oracle import pyoracle as po

    rng = np.random.default_rng(5)
    vals = np.concatenate([[0, 1, 2, 1023, 1024, 2**21 - 1], rng.integers(0, 2**21, 200)]).astype(np.uint64)
    ref = io.ref_bmi(vals)
    if ref is None:
        pytest.skip("oracle/_ref/bmi_ref is not built (needs /root/reference)")
    assert [lb.expand_bits_u64_2(int(v)) for v in vals] == [int(r) for r in ref[:, 1]]
    assert np.array_equal(io.ref_bmi(ref[:, 1])[:, 3], vals)  # contract_bits undoes expand_bits
    for bits, nbit, col in ((32, 10, 0), (64, 21, 1)):
        n = 1 << nbit
        ijk = rng.integers(0, n, size=(300, 3))
        ijk[:4] = [[0, 0, 0], [n - 1, n - 1, n - 1], [n - 1, 0, 0], [0, 0, n - 1]]
        xyz = (ijk + 0.5) / n  # cell centres of the unit box
        codes = po.morton_codes(xyz, (0, 0, 0), (1, 1, 1), len(xyz), bits=bits)
        ex = [io.ref_bmi(ijk[:, c])[:, col] for c in range(3)]
        assert np.array_equal(codes.astype(np.uint64), ex[0] * np.uint64(4) + ex[1] * np.uint64(2) + ex[2])
