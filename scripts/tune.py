#!/usr/bin/env python
"""Tuning run (under gpurun): one model at the bench size, stage times for a list of environment settings.

    python scripts/tune.py [--npart P] [--steps K] VAR=a,b,c [VAR2=x,y]

Every listed value of every variable is measured against the default (the variables are read by the
library at each launch, so one process and one setup serve all variants)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--npart", type=int, default=16 * 2**20)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--kernel", default="M4")
    ap.add_argument("--reorder", action="store_true", help="Morton-reorder the patch data after the setup")
    ap.add_argument("vars", nargs="*")
    args = ap.parse_args()
    import torch

    import bench
    from shamrock_b200 import _capi
    from tests import scenarios as S

    torch.cuda.set_device(0)
    sc = bench.workload(args.npart, 1, kernel=args.kernel)
    ctx = _capi.Context(0)
    m = S.make_cuda(sc, ctx=ctx, keep_step_data=False, fp_mode="fast")
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
    if args.reorder:
        m.reorder_particles()
    m.evolve_once()

    def measure(tag):
        for _ in range(2):
            m.set_next_dt(0.0)
            m.evolve_once()
        acc = {}
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            m.set_next_dt(0.0)
            m.evolve_once()
            for k, v in m.stage_times().items():
                acc[k] = acc.get(k, 0.0) + v / args.steps
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({"tag": tag, "ms_per_step": round(ms, 3), **{k: round(v, 3) for k, v in acc.items()}}),
              flush=True)

    measure("default")
    for spec in args.vars:
        if "+" in spec:  # a combination: A=1+B=2 sets both, one measurement
            pairs = [kv.split("=") for kv in spec.split("+")]
            for k, v in pairs:
                os.environ[k] = v
            measure(spec)
            for k, _ in pairs:
                del os.environ[k]
            continue
        name, vals = spec.split("=")
        for v in vals.split(","):
            os.environ[name] = v
            measure(f"{name}={v}")
        del os.environ[name]
    measure("default-again")


if __name__ == "__main__":
    main()
