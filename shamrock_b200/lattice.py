"""HCP lattice helpers (setup side, host): the initial-condition generator used by the reference's
SPH scripts.  Restates shammath::LatticeHCP (/root/reference/src/shammath/include/shammath/
crystalLattice.hpp:52-290,387-402) and modules::GeneratorLatticeHCP
(src/shammodels/sph/include/shammodels/sph/modules/setup/GeneratorLatticeHCP.hpp:39-120):
lattice points of the index box, x index fastest, filtered by lower <= r < upper, hpart = dr."""
import math

import numpy as np


def get_box_index_bounds(dr, box_min, box_max):
    """crystalLattice.hpp:140-163"""
    cmin = [box_min[0] / 2.0, box_min[1] / math.sqrt(3.0), box_min[2] / (2 * math.sqrt(6.0) / 3)]
    cmax = [box_max[0] / 2.0, box_max[1] / math.sqrt(3.0), box_max[2] / (2 * math.sqrt(6.0) / 3)]
    cmin = [c / dr for c in cmin]
    cmax = [c / dr for c in cmax]
    imin = [int(c) - 1 for c in cmin]  # i32(x): truncation toward zero
    imax = [int(c) + 1 for c in cmax]
    return imin, imax


def nearest_periodic_box_indices(imin, imax):
    """crystalLattice.hpp:172-198"""
    omax = list(imax)
    if imax[0] - imin[0] < 2:
        omax[0] += 1
    if (imax[1] + imin[1]) % 2 != 0:
        omax[1] += 1
    if (imax[2] + imin[2]) % 2 != 0:
        omax[2] += 1
    return list(imin), omax


def get_periodic_box(dr, imin, imax):
    """crystalLattice.hpp:108-138"""
    if imax[0] - imin[0] < 2 or (imax[1] + imin[1]) % 2 != 0 or (imax[2] + imin[2]) % 2 != 0:
        raise ValueError("x axis count should be greater than 1\ny axis count should be even\n"
                         "z axis count should be even")
    lo = [2 * imin[0], math.sqrt(3.0) * imin[1], 2 * math.sqrt(6.0) * imin[2] / 3]
    hi = [2 * imax[0], math.sqrt(3.0) * imax[1], 2 * math.sqrt(6.0) * imax[2] / 3]
    return tuple(v * dr for v in lo), tuple(v * dr for v in hi)


def get_ideal_hcp_box(dr, box_min, box_max):
    """shamrock.math.get_ideal_hcp_box (crystalLattice.hpp:387-402)"""
    imin, imax = get_box_index_bounds(dr, box_min, box_max)
    pmin, pmax = nearest_periodic_box_indices(imin, imax)
    return get_periodic_box(dr, pmin, pmax)


def hcp_positions(dr, box_min, box_max):
    """All lattice points r with box_min <= r < box_max, in the reference Iterator order
    (x index fastest, then y, then z; crystalLattice.hpp:244-248)."""
    imin, imax = get_box_index_bounds(dr, box_min, box_max)
    i = np.arange(imin[0], imax[0], dtype=np.int64)
    j = np.arange(imin[1], imax[1], dtype=np.int64)
    k = np.arange(imin[2], imax[2], dtype=np.int64)
    K, J, I = np.meshgrid(k, j, i, indexing="ij")  # x fastest in C order
    I = I.ravel()
    J = J.ravel()
    K = K.ravel()
    # generator (crystalLattice.hpp:68-80)
    x = (2 * I + (np.abs(J + K) % 2)).astype(np.float64)
    y = math.sqrt(3.0) * (J + (1.0 / 3.0) * (np.abs(K) % 2))
    z = 2 * math.sqrt(6.0) * K / 3
    r = np.stack([x, y, z], axis=1) * dr
    keep = np.ones(len(r), dtype=bool)
    for c in range(3):
        keep &= (box_min[c] <= r[:, c]) & (r[:, c] < box_max[c])
    return np.ascontiguousarray(r[keep])
