"""Patch scheduler on one GPU (shamrock_b200/csrc/scheduler.cu): split / merge of patches against the oracle's
restatement of the reference's patch-list operations (PatchCoord.hpp:36-120, scheduler_patch_list.cpp:109-185,
SchedulerPatchData.cpp:302-420) — ids, list order, data partition and the steps that follow, bit for bit — and
the automatic scheduler_step (crit_split / crit_merge).  The two-rank side (migration over NCCL, Hilbert load
balancing) is tests/test_gpu_multirank.py::scheduler."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from tests import scenarios as S  # noqa: E402
from tests.test_gpu_step import compare  # noqa: E402


def ids(m):
    return [m.patch_info(ip)["id"] for ip in range(m.patch_count)]


@pytest.mark.parametrize("fp_mode", ["strict", "fast"])
def test_forced_split_and_merge_match_the_oracle(fp_mode):
    sc = S.periodic_box(12000, "M4", "cd10", jitter=0.2, grid=(2, 1, 1))
    o, m = S.make_oracle(sc), S.make_cuda(sc, fp_mode=fp_mode)
    rtol = 0.0 if fp_mode == "strict" else 1e-10
    o.evolve_once(), m.evolve_once()
    o.split_patch(1), m.split_patch(1)
    assert m.patch_count == o.patch_count == 9
    assert ids(m) == [o.patch_id(ip) for ip in range(9)] == [0, 1, 2, 3, 4, 5, 6, 7, 8]
    for ip in range(9):
        assert m.patch_size(ip) == o.patch_size(ip)
        if fp_mode == "strict":
            assert np.array_equal(m.get(ip, "xyz"), o.get(ip, "xyz")) and np.array_equal(m.get(ip, "uint"), o.get(ip, "uint"))
    for _ in range(2):
        so, sm = o.evolve_once(), m.evolve_once()
        assert so["npart"] == sm["npart"]
        compare(m, o, sc, rtol, ints_exact=fp_mode == "strict")
    o.split_patch(3), m.split_patch(3)  # a second level: a child of the split patch
    assert ids(m) == [o.patch_id(ip) for ip in range(o.patch_count)]
    so, sm = o.evolve_once(), m.evolve_once()
    compare(m, o, sc, rtol, ints_exact=fp_mode == "strict")
    o.merge_patches(3), m.merge_patches(3)
    o.merge_patches(1), m.merge_patches(1)
    assert m.patch_count == o.patch_count == 2 and ids(m) == [0, 1]
    so, sm = o.evolve_once(), m.evolve_once()
    compare(m, o, sc, rtol, ints_exact=fp_mode == "strict")


def test_merge_needs_a_complete_octet():
    from shamrock_b200 import _capi

    sc = S.periodic_box(4000, "M4", "cd10", grid=(2, 1, 1))
    m = S.make_cuda(sc)
    with pytest.raises(_capi.ShamB200Error, match="octet"):
        m.merge_patches(0)


def test_automatic_scheduler_step():
    """crit_split / crit_merge: a patch above crit_split is split into its eight children at the start of the step,
    an octet below crit_merge is merged back; the object count is conserved and the physics goes on"""
    sc = S.periodic_box(16000, "M4", "cd10", jitter=0.1, grid=(1, 1, 1))
    m = S.make_cuda(sc, fp_mode="fast", keep_step_data=False)
    n = len(sc["xyz"])
    m.init_scheduler(n // 2, 1, step_freq=1)  # 16 k objects in one patch > 8 k: split; children ~2 k: stay
    st = m.evolve_once()
    log = m.scheduler_log()
    assert log["splits"] == 1 and m.patch_count == 8 and st["npart"] == n
    assert sum(m.patch_size(ip) for ip in range(8)) == n
    st = m.evolve_once()
    assert m.scheduler_log()["splits"] == 0 and m.patch_count == 8
    m.init_scheduler(10 * n, 2 * n, step_freq=1)  # now everything is below crit_merge: merged back
    st = m.evolve_once()
    assert m.scheduler_log()["merges"] == 1 and m.patch_count == 1 and m.patch_size(0) == n
    assert np.isfinite(m.get(0, "axyz")).all() and st["npart"] == n


def test_many_patches():
    """two levels of splits of a 2 x 2 x 2 grid: 64 + ... patches (the round-1 limit was 256 patches in total)"""
    sc = S.periodic_box(40000, "M4", "cd10", jitter=0.1, grid=(2, 2, 2))
    m = S.make_cuda(sc, fp_mode="fast", keep_step_data=False)
    n = len(sc["xyz"])
    for _ in range(2):
        for ip in range(m.patch_count):
            m.split_patch(ip)
    assert m.patch_count == 512
    st = m.evolve_once()
    st = m.evolve_once()
    assert st["npart"] == n and sum(m.patch_size(ip) for ip in range(m.patch_count)) == n
