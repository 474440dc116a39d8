#!/usr/bin/env python
"""Host-resident step at the bench size (under gpurun): wall time per shamb200_model_evolve_once_host and the stage
times of its kernels for a list of environment settings (read by the library at every call).

    python scripts/e2e_probe.py [--npart P] [--steps K] SHAMB200_HOST_SLICES=1,2,4,8
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--npart", type=int, default=16 * 2**20)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("vars", nargs="*")
    args = ap.parse_args()
    import torch

    import bench
    from shamrock_b200 import _capi
    from tests import scenarios as S

    torch.cuda.set_device(0)
    ctx = _capi.Context(0)
    sc = S.periodic_box(args.npart, "M4", "cd10", jitter=0.0, sort_mode="radix", local_boxes="device")
    m = S.make_cuda(sc, ctx=ctx, keep_step_data=False, fp_mode="fast")
    S.periodic_box_on_device(m, args.npart, "M4")
    m.reorder_particles()
    m.evolve_once()
    for _ in range(2):
        m.set_next_dt(0.0)
        m.evolve_once()
    IN = ["xyz", "vxyz", "axyz", "hpart", "uint", "duint", "alpha_AV", "soundspeed"]
    OUT = [nm for nm, _ in bench.MAIN_FIELDS if nm != "axyz_ext"]
    host = {}
    for nm, nv in bench.MAIN_FIELDS:
        a = m.get(0, nm)
        t = torch.empty(a.size, dtype=torch.float64).pin_memory()
        t.numpy()[:] = a.reshape(-1)
        host[nm] = t
    n = m.patch_size(0)

    def step():
        m.set_next_dt(0.0)
        m.evolve_once_host(0, n, {nm: host[nm].data_ptr() for nm in IN}, {nm: host[nm].data_ptr() for nm in OUT})

    def measure(tag):
        step()
        acc, t = {}, []
        for _ in range(args.steps):
            ctx.synchronize()
            t0 = time.perf_counter()
            step()
            t.append((time.perf_counter() - t0) * 1e3)
            for k, v in m.stage_times().items():
                acc[k] = acc.get(k, 0.0) + v / args.steps
        print(json.dumps({"tag": tag, "wall_ms": [round(x, 2) for x in t], "id_ranges": m.host_step_info(0)[0],
                          "kernels_ms": round(sum(acc.values()), 2), **{k: round(v, 2) for k, v in acc.items()}}),
              flush=True)

    measure("default")
    for spec in args.vars:
        name, vals = spec.split("=")
        for v in vals.split(","):
            os.environ[name] = v
            measure(f"{name}={v}")
        del os.environ[name]


if __name__ == "__main__":
    main()
