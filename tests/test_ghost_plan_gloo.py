"""Host-side logic of the multi-GPU path on CPU: two processes (gloo, world_size 2) run the patch→rank
deal and the ghost-interface planner of libshamb200.so (host-only entry points), exchange the ghost
positions in the planned order, and each rank checks the merged {xyz, h} of its own patches
bit-for-bit against the single-process oracle (BasicSPHGhostHandler semantics: block order
(sender, receiver), periodic images x→y→z, ids ascending inside a block)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from tests import scenarios as S  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, grid, periodic, q, balanced=False):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from shamrock_b200 import _capi

        sc = S.periodic_box(5000, "M4", "cd10", jitter=0.2, grid=grid) if periodic else S.disc(4000, "M4", grid=grid)
        sc["cfg"]["h_max_subcycles_count"] = 1  # the ghost zones of the FIRST sub-cycle (initial h)
        o = S.make_oracle(sc)  # single-process truth (both ranks build it; it is the checker)
        st = o.evolve_once()  # first step: dt = 0, positions only see the boundary wrap
        assert st["time"] == 0.0
        boxes, owner = _capi.plan_patch_grid(sc["bmin"], sc["bmax"], grid, world)
        npatch = len(boxes)
        assert sorted(set(owner.tolist())) == list(range(world))  # every rank owns patches
        assert (np.diff(owner) >= 0).all()  # contiguous deal of patch ids
        if balanced:  # owners from the Hilbert-curve load balancer on the particle counts of the setup
            G = 1 << 21
            coords = np.array([(x * (G // grid[0]), y * (G // grid[1]), z * (G // grid[2]))
                               for z in range(grid[2]) for y in range(grid[1]) for x in range(grid[0])], dtype=np.uint64)
            loads = S.patch_loads(sc, world)
            owner, _ = _capi.plan_load_balance(coords, loads, world)
            t = torch.from_numpy(owner.astype(np.int64))
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # every rank computed the same table
            assert np.array_equal(t.numpy(), owner)
            per = np.bincount(owner, weights=loads.astype(np.float64), minlength=world)
            assert per.min() > 0 and per.max() <= 0.75 * loads.sum()
        # local patch data = the particles of my patches, in push order (the oracle's pre-step order)
        xyz, h = sc["xyz"], sc["hpart"]
        if periodic:  # apply_position_boundary runs before the ghost exchange (integrators.cpp:207-236)
            lo = np.array(sc["bmin"])
            d = np.array(sc["bmax"]) - lo
            xyz = np.fmod(np.fmod(xyz - lo, d) + d, d) + lo
        inside = lambda k: np.all((boxes[k, 0] <= xyz) & (xyz < boxes[k, 1]), axis=1)  # noqa: E731
        mine = [k for k in range(npatch) if owner[k] == rank]
        local = {k: (xyz[inside(k)], h[inside(k)]) for k in mine}
        if not periodic:  # accretion + kill sphere happen before the ghost exchange in the step
            for k in mine:
                x, hh = local[k]
                r2 = (x * x).sum(1)
                keep = (r2 > sc["cfg"]["pm_racc"] ** 2) & ~(np.sqrt(r2) > sc["kill"][0][1])
                local[k] = (x[keep], hh[keep])
        Rk = S.RKERN[sc["kernel"]]
        meta = torch.zeros(2 * npatch, dtype=torch.float64)
        for k in mine:
            if len(local[k][1]):
                meta[k] = local[k][1].max()
                meta[npatch + k] = len(local[k][1])
        dist.all_reduce(meta, op=dist.ReduceOp.MAX)  # C-ABI model does the same over NCCL
        pcount = meta[npatch:].numpy().astype(np.uint32)
        interact = np.where(pcount > 0, meta[:npatch].numpy() * 1.1 * Rk, -np.finfo(np.float64).max)
        itfs = _capi.plan_interfaces(boxes.reshape(npatch, 6), sc["bmin"], sc["bmax"], periodic, interact, pcount)
        # ghost ids of my senders + counts for everybody
        ids, counts = {}, torch.zeros(len(itfs), dtype=torch.int64)
        for q_, it in enumerate(itfs):
            if owner[it.sender] == rank:
                x = local[it.sender][0]
                sel = np.all((np.array(it.cut_lo) <= x) & (x < np.array(it.cut_hi)), axis=1)
                ids[q_] = np.nonzero(sel)[0]
                counts[q_] = len(ids[q_])
        dist.all_reduce(counts, op=dist.ReduceOp.MAX)
        ghosts = {k: [] for k in mine}
        reqs, recv_bufs = [], []
        for q_, it in enumerate(itfs):  # same order on every rank → matching send/recv pairs
            c = int(counts[q_])
            if c == 0:
                continue
            so, ro = owner[it.sender], owner[it.receiver]
            if so == rank:
                x, hh = local[it.sender]
                blk = np.concatenate([x[ids[q_]] + np.array(it.offset), hh[ids[q_]][:, None]], axis=1)
                if ro == rank:
                    ghosts[it.receiver].append(blk)
                else:
                    reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(blk)), dst=int(ro), tag=q_))
            elif ro == rank:
                buf = torch.empty((c, 4), dtype=torch.float64)
                reqs.append(dist.irecv(buf, src=int(so), tag=q_))
                ghosts[it.receiver].append(buf)
        for r in reqs:
            r.wait()
        for k in mine:
            if not len(local[k][1]):
                continue
            g = [b.numpy() if isinstance(b, torch.Tensor) else b for b in ghosts[k]]
            merged = np.concatenate([np.concatenate([local[k][0], local[k][1][:, None]], axis=1)] + g)
            ref_xyz, ref_h = o.get(k, "step.mxyz"), o.get(k, "step.mh")
            assert merged.shape[0] == len(ref_h), (k, merged.shape, len(ref_h))
            assert np.array_equal(merged[:, :3], ref_xyz), f"patch {k}: merged positions differ"
            assert np.array_equal(merged[:, 3], ref_h), f"patch {k}: merged h differ"
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc() + repr(e)))


@pytest.mark.parametrize("grid,periodic", [((2, 1, 1), True), ((2, 2, 1), True), ((2, 2, 2), True), ((2, 2, 1), False)])
def test_ghost_exchange_plan_world2(grid, periodic):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, grid, periodic, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_ghost_exchange_with_balanced_owner_table_world2():
    """patch -> rank from shamb200_plan_load_balance (not contiguous in patch id): the planned exchange still
    reproduces the oracle's merged ghost zones"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, (4, 2, 2), False, q, True)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_patch_grid_matches_reference_coordinates():
    from shamrock_b200 import _capi

    boxes, owner = _capi.plan_patch_grid((-1.0, -2.0, 0.0), (1.0, 2.0, 8.0), (4, 2, 1), 4)
    assert boxes.shape == (8, 2, 3) and owner.tolist() == [0, 0, 1, 1, 2, 2, 3, 3]
    assert np.array_equal(boxes[0, 0], [-1.0, -2.0, 0.0]) and np.array_equal(boxes[-1, 1], [1.0, 2.0, 8.0])
    assert np.array_equal(boxes[1, 0], [-0.5, -2.0, 0.0])  # x fastest
    with pytest.raises(_capi.ShamB200Error, match="powers of two"):
        _capi.plan_patch_grid((0, 0, 0), (1, 1, 1), (3, 1, 1), 1)
