#!/usr/bin/env python
"""Runs one of the BASELINE.json configurations on one GPU for a few real timesteps and prints the rate
(stress / capacity check at sizes the parity tests do not reach; parity itself is tests/).

    python scripts/run_config.py --config C4|C3|C5|C1 [--npart N] [--steps K] [--fp fast|strict]

C1: Sod tube (sod_tube_sph.py geometry, M4 as BASELINE.json names it), C3: periodic Sedov-like box with the
M6 kernel, C5: disc around a point mass (free boundaries, LP07, ConstantDisc viscosity, accretion, kill sphere).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import scenarios as S  # noqa: E402
from tests import sod_tube  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3")
    ap.add_argument("--npart", type=int, default=8 * 2**20)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--fp", default="fast")
    ap.add_argument("--mc", action="store_true", help="C5: Monte-Carlo positions instead of the regular lattice")
    a = ap.parse_args()
    if a.config == "C4":  # the bench workload with REAL timesteps (bench.py replays dt = 0)
        sc = S.periodic_box(a.npart, "M4", "cd10", sort_mode="radix")
    elif a.config == "C3":
        sc = S.periodic_box(a.npart, "M6", "cd10", sort_mode="radix")
    elif a.config == "C5":
        sc = S.disc(a.npart, "M4", sort_mode="radix", regular=not a.mc)
    elif a.config == "C1":
        sc = sod_tube.scenario(kernel="M4")
        sc["sort_mode"] = "radix"
    else:
        raise SystemExit("unknown config")
    m = S.make_cuda(sc, keep_step_data=False, fp_mode=a.fp)
    st = m.evolve_once()  # dt = 0: converges h, first forces
    out, tols = [], []
    for _ in range(a.steps):
        t0 = time.perf_counter()
        st = m.evolve_once()
        out.append((time.perf_counter() - t0) * 1e3)
        tols.append(m.list_tolerance())
    stages = m.stage_times()
    print(json.dumps({"config": a.config, "scenario": sc["name"], "npart": int(st["npart"]), "fp": a.fp,
                      "ms_per_step": out, "best_part_per_s": st["npart"] / (min(out) * 1e-3),
                      "neighbours_per_particle": st["K_local"] / max(st["n_local"], 1), "time": st["time"],
                      "dt": st["dt"], "h_subcycles": st["h_subcycles"], "h_iters": st["h_iters_last"],
                      "corrector_iter": st["corrector_iter"],
                      # neighbour-list tolerance of every step (the reference always uses 1.1), the h growth the step
                      # needed, and how many steps had to be redone with 1.1 (shamb200_model_list_tolerance)
                      "list_tolerance_per_step": [round(t["last"], 4) for t in tols],
                      "h_growth_per_step": [round(t["growth"], 5) for t in tols],
                      "list_fallbacks": tols[-1]["fallbacks"] if tols else 0,
                      "stage_ms_last": {k: round(v, 3) for k, v in stages.items()}}))


if __name__ == "__main__":
    main()
