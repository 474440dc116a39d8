"""Host-side setup helpers: the rank-local particle generation used by `bench.py --gpus N` gives, rank by
rank, exactly the particles of the single-process setup (positions, internal energy, particle mass)."""
import numpy as np
import pytest

from shamrock_b200 import _capi
from tests import scenarios as S


@pytest.mark.parametrize("world", [2, 4])
def test_rank_local_generation_matches_global(world):
    kw = dict(grid=(world, 1, 1), stretch=(world, 1, 1), sort_mode="radix")
    g = S.periodic_box(6000 * world, "M4", "cd10", **kw)
    tot = len(g["xyz"])
    xyz, u = [], []
    for rank in range(world):
        def local_boxes(bmin, bmax, rank=rank):
            boxes, owner = _capi.plan_patch_grid(bmin, bmax, (world, 1, 1), world)
            return [(boxes[k][0], boxes[k][1]) for k in range(len(owner)) if owner[k] == rank]

        sc = S.periodic_box(6000 * world, "M4", "cd10", local_boxes=local_boxes, count_reduce=lambda n: tot, **kw)
        assert sc["cfg"]["gpart_mass"] == g["cfg"]["gpart_mass"] and sc["bmin"] == g["bmin"]
        assert len(sc["xyz"]) > 0
        xyz.append(sc["xyz"])
        u.append(sc["uint"])
    xyz, u = np.concatenate(xyz), np.concatenate(u)
    assert len(xyz) == tot
    ka, kg = np.lexsort(xyz.T[::-1]), np.lexsort(g["xyz"].T[::-1])
    assert np.array_equal(xyz[ka], g["xyz"][kg]) and np.array_equal(u[ka], g["uint"][kg])
