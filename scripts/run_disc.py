#!/usr/bin/env python
"""BASELINE.json config C5: the protoplanetary disc (central point mass, free boundaries, locally isothermal
LP07 EOS, ConstantDisc viscosity, accretion radius, kill sphere; examples/sph/run_circular_disc_central_pot.py)
at 64 Mi particles on 8 GPUs — or any size / rank count:

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_disc.py \
        --npart 67108864 --steps 5

Everything runs on the devices: the Monte-Carlo disc is generated patch by patch by the rank that owns the
patch (shamb200_model_add_disc_mc, counter-based draws), the patch scheduler splits the dense patches
(crit_split) and deals the patches to the ranks along the Hilbert curve by particle count, moving whole patches
over NCCL, before the first step and every `--sched-freq` steps.  Prints one JSON line (rank 0): rate, per-rank
load imbalance, scheduler log."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--npart", type=int, default=64 * 2**20)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--grid", default="8,8,2")
    ap.add_argument("--crit-split", type=int, default=0, help="0: npart / (12 * world)")
    ap.add_argument("--sched-freq", type=int, default=2)
    ap.add_argument("--fp", default="fast")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--mc", action="store_true",
                    help="Monte-Carlo disc (GeneratorMCDisc) instead of the regular lattice disc: unrelaxed random "
                         "positions leave a few objects whose h iteration never converges (100 sub-cycles per step, "
                         "as in the reference)")
    a = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist

    from shamrock_b200 import _capi
    from tests import scenarios as S

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    nccl_id = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ids = [_capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        nccl_id = ids[0]
    grid = tuple(int(v) for v in a.grid.split(","))
    sc = S.disc(1000, "M4", grid=grid, sort_mode="radix")  # configuration, box and kill sphere of the scenario
    for k in ("xyz", "vxyz", "hpart", "uint"):
        sc[k] = sc[k][:0]
    sc["kill"] = [((0.0, 0.0, 0.0), 4.0)]
    disc_mass = 0.01
    ctx = _capi.Context(local)
    t0 = time.perf_counter()
    m = S.make_cuda(sc, ctx=ctx, keep_step_data=False, rank=rank, world=world, nccl_id=nccl_id, fp_mode=a.fp)
    m.set_particle_mass(disc_mass / a.npart)
    def say(*msg):
        if rank == 0:
            print(f"[{time.perf_counter() - t0:7.2f}s]", *msg, file=sys.stderr, flush=True)

    if a.mc:
        n = m.add_disc_mc(a.npart, a.seed, 1.0, 3.0, 1.0, 0.25, 0.05, disc_mass)
    else:  # tests/scenarios.disc(regular=True) on the device: HCP lattice in the flared volume
        import math

        rin, rout, H_r = 1.0, 3.0, 0.08
        vol = math.pi * (rout**2 - rin**2) * 3 * H_r * 2.0
        dr = (vol / (a.npart * 4 * math.sqrt(2))) ** (1.0 / 3.0)
        n = m.add_disc_lattice(dr, rin * 1.05, rout, 1.5 * H_r)
        m.set_particle_mass(disc_mass / n)
    say("disc generated", n)
    big = ([-1e300] * 3, [1e300] * 3)
    m.set_value_in_a_box("uint", 1e-3, *big)
    crit_split = a.crit_split or max(1, a.npart // (12 * world))
    m.init_scheduler(crit_split, 1, step_freq=a.sched_freq)
    sched = []
    for _ in range(6):  # split until every patch is below crit_split, balance after every round
        log = m.scheduler_step(True, True)
        sched.append(log)
        say("scheduler", log)
        if not log["splits"]:
            break
    m.reorder_particles()
    ctx.synchronize()
    t_setup = time.perf_counter() - t0
    say("setup done")
    st = m.evolve_once()  # dt = 0: converges h, first forces
    say("first step", st)
    ms = []
    for _ in range(a.steps):
        if world > 1:
            dist.barrier()
        t1 = time.perf_counter()
        st = m.evolve_once()
        ctx.synchronize()
        dt_ms = (time.perf_counter() - t1) * 1e3
        if world > 1:
            t = torch.tensor([dt_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_ms = float(t.item())
        ms.append(dt_ms)
    n_loc = sum(m.patch_size(ip) for ip in range(m.patch_count) if m.patch_is_local(ip))
    loads = [n_loc]
    if world > 1:
        loads = [None] * world
        dist.all_gather_object(loads, n_loc)
    if rank == 0:
        stages = m.stage_times()
        mean = sum(loads) / len(loads)
        print(json.dumps({
            "config": "C5 disc", "npart": int(st["npart"]), "generated": n, "n_gpus": world, "fp": a.fp,
            "patches": m.patch_count, "crit_split": crit_split, "setup_s": round(t_setup, 2),
            "ms_per_step": [round(v, 2) for v in ms], "best_part_per_s": st["npart"] / (min(ms) * 1e-3),
            "rank_loads": loads, "imbalance": max(loads) / mean - 1.0, "scheduler_setup": sched,
            "scheduler_last": m.scheduler_log(), "h_subcycles": st["h_subcycles"], "h_iters": st["h_iters_last"],
            "corrector_iter": st["corrector_iter"], "dt": st["dt"], "time": st["time"],
            "neighbours_per_particle": st["K_local"] / max(st["n_local"], 1),
            "stage_ms_last": {k: round(v, 3) for k, v in stages.items()}}), flush=True)
    m.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
