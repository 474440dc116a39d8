#!/usr/bin/env python
"""Runs the reference's Sod-tube CI case (examples/tests_ci/sod_tube_sph.py) to t = 0.245 with the CPU
oracle (default) or the CUDA path and compares the L2 distances with the constants of the reference's
script.  Writes tests/golden/sod_tube_<impl>.json (committed: the oracle run takes minutes).

    python tests/golden/run_sod_known_answer.py [oracle|cuda-strict|cuda-fast] [--niter N]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import scenarios as S  # noqa: E402
from tests import sod_tube as sod  # noqa: E402


def main():
    impl = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "oracle"
    niter = int(sys.argv[sys.argv.index("--niter") + 1]) if "--niter" in sys.argv else -1
    sc = sod.scenario()
    print(f"{sc['name']}: N = {len(sc['xyz'])}, pmass = {sc['cfg']['gpart_mass']!r}", flush=True)
    if impl == "oracle":
        m = S.make_oracle(sc)
    else:
        m = S.make_cuda(sc, fp_mode=impl.split("-")[1], keep_step_data=False)
    t0 = time.time()
    # fingerprint of the state after 3 iterations: lets a quick test check that a committed result file
    # belongs to the code that is in the tree (tests/test_sod_known_answer.py)
    sod.evolve_until(m, sod.T_TARGET, 3)
    checkpoint = sod.fingerprint(m)
    if niter >= 0:
        niter = max(niter - 3, 0)
    n = 3 + (sod.evolve_until(m, sod.T_TARGET, niter,
                              log=lambda s: print(s, f"[{time.time() - t0:.0f}s]", flush=True)) if niter != 0 else 0)
    st = m.state()
    l2 = sod.compute_L2_dist(m.get(0, "xyz"), m.get(0, "vxyz"), m.get(0, "hpart"), m.get(0, "uint"),
                             sc["cfg"]["gpart_mass"], S.HFACT[sc["kernel"]])
    rel = sod.relative_errors(l2)
    out = {"impl": impl, "iterations": n, "time": st["time"], "npart": len(sc["xyz"]), "L2": l2,
           "expected": sod.EXPECTED, "relative_error": rel, "checkpoint_iter3": checkpoint, "wall_s": time.time() - t0}
    print(json.dumps(out, indent=1))
    if niter < 0:
        with open(os.path.join(ROOT, "tests", "golden", f"sod_tube_{impl.replace('-', '_')}.json"), "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
