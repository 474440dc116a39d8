#!/usr/bin/env python
"""Tree / neighbour micro-bench (BASELINE.json config C2): Morton codes + radix sort + leaf compression +
Karras LBVH + AABBs (shamb200_tree_build), per-cell max h (shamb200_tree_field_max) and the neighbour cache
(shamb200_neigh_cache_build) on one B200, through the stage-level C ABI on device pointers.

    python scripts/bench_tree.py [--sizes 1,4,16,64,256] [--cache-max 16] [--reps 5]

Positions: HCP lattice in the unit cube (generated on the device, x fastest), h = 1.2 (m/rho)^(1/3) (M4).  Prints one
JSON line per size: particles/s of the tree build and of tree + cache, and the achieved fraction of the HBM
bound of SURVEY.md §8d (176 M + 90 L bytes for the tree; + 64 M + 4 K + 12 N for the cache).
"""
import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shamrock_b200 import _capi  # noqa: E402


def hcp_device(n_target):
    """HCP lattice points of the unit cube, generated on the device (crystalLattice.hpp:68-80)"""
    dr = (1.0 / (n_target * 4 * math.sqrt(2))) ** (1.0 / 3.0)
    ni = int(1.0 / (2 * dr)) + 1
    nj = int(1.0 / (math.sqrt(3.0) * dr)) + 1
    nk = int(1.0 / (2 * math.sqrt(6.0) / 3 * dr)) + 1
    k, j, i = torch.meshgrid(torch.arange(nk, device="cuda"), torch.arange(nj, device="cuda"),
                             torch.arange(ni, device="cuda"), indexing="ij")
    i, j, k = i.reshape(-1).double(), j.reshape(-1).double(), k.reshape(-1).double()
    x = (2 * i + torch.remainder(j + k, 2)) * dr
    y = math.sqrt(3.0) * (j + torch.remainder(k, 2) / 3.0) * dr
    z = 2 * math.sqrt(6.0) / 3 * k * dr
    keep = (x < 1) & (y < 1) & (z < 1)
    xyz = torch.stack([x[keep], y[keep], z[keep]], dim=1).contiguous()
    return xyz, dr


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1,4,16,64,256", help="millions (2^20) of particles")
    ap.add_argument("--cache-max", type=float, default=16, help="largest size (Mi) that also builds the cache")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
    # the library works on a torch stream so that torch events (recorded on the current stream) time it
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    ctx = _capi.Context(0, stream=st.cuda_stream)
    for s in [float(v) for v in a.sizes.split(",")]:
        xyz, dr = hcp_device(int(s * 2**20))
        n = xyz.shape[0]
        # h = hfact (m / rho)^(1/3) with 4 sqrt(2) dr^3 of volume per particle (M4: ~69 list entries each)
        h = torch.full((n,), 1.2 * (4 * math.sqrt(2)) ** (1.0 / 3.0) * dr, dtype=torch.float64, device="cuda")
        state = {}

        def tree():
            state["tv"] = ctx.tree_build(xyz, n, [0, 0, 0], [1, 1, 1], reduction_level=3, sort_mode="radix")

        ms_tree = timed(tree, a.reps)
        tv = state["tv"]
        L = tv.leaf_count
        line = {"npart": n, "leaves": L, "tree_ms": ms_tree, "tree_part_per_s": n / (ms_tree * 1e-3),
                "tree_frac_of_hbm_bound": (176 * n + 90 * L) / (hbm * 1e9) / (ms_tree * 1e-3)}
        if s <= a.cache_max:
            rint = torch.empty(tv.leaf_count + tv.int_count, dtype=torch.float64, device="cuda")

            def cache():
                tree()
                ctx.tree_field_max(state["tv"], h, 1.1, rint)
                state["cv"] = ctx.neigh_cache_build(state["tv"], xyz, h, rint, n, 2.0, 1.1, True)

            ms_all = timed(cache, max(2, a.reps // 2))
            K = state["cv"].sum_neigh_cnt
            line.update({"tree_plus_cache_ms": ms_all, "tree_plus_cache_part_per_s": n / (ms_all * 1e-3),
                         "neighbours_per_particle": K / n,
                         "cache_frac_of_hbm_bound": (176 * n + 90 * L + 64 * n + 4 * K + 12 * n) / (hbm * 1e9)
                         / (ms_all * 1e-3)})
        print(json.dumps(line), flush=True)
        del xyz, h
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
