#!/usr/bin/env python
"""bench.py — full SPH step throughput (particle-updates/s) of the B200 path, reference protocol.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--npart-per-gpu P] [--impl reference]

Workload (BASELINE.json config C4, the one the metric is quoted on at 1/2/4/8 GPUs; protocol of the
reference's examples/benchmarks/sph_homogeneous_benchmark.py): HCP lattice in a periodic box, M4 kernel,
CD10 artificial viscosity, adiabatic gamma = 5/3, Sedov-like `uint` injection, C_cour = C_force = 0.1,
16 Mi particles per GPU (weak scaling: the box is stretched along x, one slab of patches per GPU);
W warm-up timestep()s, then K x { set_next_dt(0); timestep() }.  rate = sum_ranks N / max_ranks t.

One JSON line on stdout (rank 0).  `value`: patch data resident in HBM.  `e2e`: the same step driven
through the C ABI with HOST patch data (shamb200_model_evolve_once_host): every step uploads the step's
input fields from pinned host memory and reads all 12 main-layout fields back.
`--impl reference`: the CPU oracle (a port of the reference's algorithms; the reference itself needs
SYCL + MPI and cannot be built in this image) on all host cores, on a bounded sample of the workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H100_PUBLISHED = 25.50e6  # reference full-step rate, 1x H100, same protocol (N = 33.8 M)
MAIN_FIELDS = [("xyz", 3), ("vxyz", 3), ("axyz", 3), ("axyz_ext", 3), ("hpart", 1), ("uint", 1), ("duint", 1),
               ("alpha_AV", 1), ("divv", 1), ("dtdivv", 1), ("curlv", 3), ("soundspeed", 1)]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None  # host time window of the timed region

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        # nvidia-smi needs ~1 s to start: it is launched before the warm-up and only the samples taken
        # between begin() and end() (the timed region) count
        rows = [r[1:] for r in self.rows if self.t0 is None or self.t0 <= r[0] <= (self.t1 or r[0])]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def bind_to_gpu_numa_node(index):
    """Run this process on the CPUs next to its GPU (NVML affinity mask) before any page-locked host memory
    is allocated: the host patch data of the e2e leg then sits on the NUMA node the GPU's PCIe link hangs
    off, instead of wherever the launcher started the process."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        old = os.sched_getaffinity(0)
        cpus &= old
        if cpus:
            os.sched_setaffinity(0, cpus)
        return old
    except Exception:
        return None


def workload(npart_per_gpu, n_gpus, kernel="M4", rank=0, count_reduce=None):
    from tests import scenarios as S

    # weak scaling: box stretched along x, patch grid = n_gpus slabs (power of two).  With several ranks
    # each one generates only the particles of its own patches (the global count is all-reduced)
    local_boxes = None
    if n_gpus > 1 and count_reduce is not None:
        from shamrock_b200 import _capi

        def local_boxes(bmin, bmax):
            boxes, owner = _capi.plan_patch_grid(bmin, bmax, (n_gpus, 1, 1), n_gpus)
            return [(boxes[k][0], boxes[k][1]) for k in range(len(owner)) if owner[k] == rank]

    sc = S.periodic_box(npart_per_gpu * n_gpus, kernel, "cd10", jitter=0.0, grid=(n_gpus, 1, 1),
                        stretch=(n_gpus, 1, 1), sort_mode="radix", local_boxes=local_boxes,
                        count_reduce=count_reduce)
    return sc


# algorithmic bytes and FP64 flops per launch of the heavy stages.
# M merged, N real, K = sum of list lengths, K_acc ~ K / htol^3 pairs inside the kernel support, L leaves.
# Flop convention: FMA = 2, div / sqrt = 1; candidate test 10, density pair 39, div+curl+dtdivv 175,
# force + v_sig 155 per accepted pair.
def alg_work(stage, N, M, K, L, sweeps, tests=0, omega_in_av=True, list_tol=1.1):
    """omega_in_av: fast fp + CD10 / MM97 — the Ω sum (one density-type pass, 39 flop per pair) is evaluated inside
    the CD10 operator pass instead of after the h iteration"""
    K_acc = K / list_tol**3  # K counts the lists as built: radius R h list_tol (shamb200_model_list_tolerance)
    h_passes = sweeps + (0 if omega_in_av else 1)
    return {
        # the search = one tree walk per group of 8 leaves + the accept / fill kernel; SURVEY.md §8d's figure
        # for the whole cache (64 M + 4 K + 12 N) split over the two launches: packed nodes (64 B, I + L = 2 L)
        # and candidate entries for the walk; sorted records (32 B), list and count / offset writes for the lists
        "neigh_walk": (64 * 2 * L + 8 * 12 * L, 30.0 * 2 * L),
        "neigh_lists": (32 * M + 4 * K + 8 * N, 10.0 * (tests or K)),  # one accept test per (particle, candidate)
        "h_iteration": (h_passes * 4 * K + 32 * M + 40 * N, h_passes * (10.0 * K + 39.0 * K_acc)),
        "divv_curlv_dtdivv": (4 * K + 96 * M + 40 * N, 10.0 * K + (175.0 + (39.0 if omega_in_av else 0.0)) * K_acc),
        "forces": (4 * K + 128 * M + 64 * N, 10.0 * K + 155.0 * K_acc),
        "build_trees": (176 * M + 90 * L, 0.0),
    }.get(stage)


def host_cores():
    """cores this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers: undo that for the CPU arm)"""
    try:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
    except Exception:
        pass
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def oracle_all_cores():
    """load the oracle with every host core: OMP_NUM_THREADS must be right BEFORE libgomp starts"""
    n = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    from oracle import pyoracle as po

    po.set_num_threads(n)
    return po, n


def run_reference(args):
    """CPU arm: the oracle port on all host cores, bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    po, ncores = oracle_all_cores()
    from tests import scenarios as S

    n_sample = args.cpu_sample
    sc = workload(n_sample, 1)
    o = S.make_oracle(sc)
    n = len(sc["xyz"])
    for _ in range(max(args.warmup, 1)):
        o.evolve_once()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.set_next_dt(0.0)
        o.evolve_once()
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    line = {
        "impl": "reference", "metric": "SPH particle-updates/sec (full step)", "value": v, "unit": "particles/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C4 periodic HCP box, M4, CD10 AV, adiabatic, dt=0 replay (sph_homogeneous_benchmark "
                   "protocol)", "npart_sample": n},
        "cpu_baseline": {"value": v, "unit": "particles/s", "cores": ncores, "kind": "port",
                         "sample": f"{n} particles of the same box (oracle port of the reference algorithms, "
                                   f"OpenMP), {args.steps} dt=0 replays"},
        "e2e": {"value": v, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def rel_err(a, b):
    """max over particles of |a - b| / max(|b|, mean|b|): the north-star tolerance (1e-10) is on this number"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return float("inf")
    if a.size == 0:
        return 0.0
    scale = np.maximum(np.abs(b), np.abs(b).mean())
    return float((np.abs(a - b) / np.where(scale > 0, scale, 1.0)).max())


CHECK_FIELDS = ["xyz", "vxyz", "hpart", "uint", "axyz", "duint", "alpha_AV", "divv", "dtdivv", "curlv", "soundspeed"]
INT_NAMES = ["tree.sorted_morton", "tree.sort_index_map", "tree.reduc_index_map", "tree.reduced_morton",
             "tree.lchild_id", "tree.rchild_id", "tree.endrange", "cache.cnt_neigh", "cache.index_neigh_map"]


def cpu_baseline(args, ctx):
    """the oracle timed on a bounded sample of the workload, all host cores — and, on the same sample, the
    CUDA model checked against the state the oracle ends in (not only timed)"""
    po, ncores = oracle_all_cores()
    from tests import scenarios as S

    sc = workload(args.cpu_sample, 1)
    o = S.make_oracle(sc)
    n = len(sc["xyz"])
    o.evolve_once()
    t0 = time.perf_counter()
    k = 2
    for _ in range(k):
        o.set_next_dt(0.0)
        so = o.evolve_once()
    dt = time.perf_counter() - t0
    m = S.make_cuda(sc, ctx=ctx, keep_step_data=False, fp_mode=args.fp)
    m.evolve_once()
    for _ in range(k):
        m.set_next_dt(0.0)
        sm = m.evolve_once()
    errs = {nm: rel_err(m.get(0, nm), o.get(0, nm)) for nm in CHECK_FIELDS}
    errs["dt"] = abs(sm["dt"] - so["dt"]) / abs(so["dt"])
    # the model's lists are built with the radius R h tol, tol <= the reference's 1.1 (shamb200_model_list_tolerance):
    # equal counts when tol = 1.1, otherwise every list is a subset of the reference's
    cm, co = m.get(0, "cache.cnt_neigh"), o.get(0, "cache.cnt_neigh")
    ltol = m.list_tolerance()
    cnt_equal = bool(np.array_equal(cm, co))
    cnt_subset = bool(cm.shape == co.shape and (cm <= co).all())
    m.close()
    return {"value": n * k / dt, "unit": "particles/s", "cores": ncores, "kind": "port",
            "sample": f"{n} particles of the same periodic box, 1 warm-up + {k} dt=0 replays, oracle (C++/OpenMP port "
                      "of the reference algorithms, -O3 -march=native; the SYCL reference cannot be built here)",
            "parity_on_sample": {
                "n": n, "steps": 1 + k,
                # the north star's fields (density = h, acceleration, du/dt; plus what they are integrated into)
                "max_rel_err": max(errs[f] for f in ("xyz", "vxyz", "hpart", "uint", "axyz", "duint", "dt")),
                # the CD10 switch on the UNPERTURBED lattice of the reference protocol: div v, curl v and their
                # time derivative are sums that cancel to round-off away from the blast, alpha is a ratio of them
                "max_rel_err_av_switch": max(errs[f] for f in ("alpha_AV", "divv", "dtdivv", "curlv", "soundspeed")),
                "per_field": {f: float(f"{e:.3e}") for f, e in errs.items()},
                "list_tolerance": ltol["last"], "list_fallbacks": ltol["fallbacks"],
                "neighbour_counts_equal": cnt_equal, "neighbour_counts_subset_of_reference": cnt_subset,
                "neighbours_per_particle": [float(cm.mean()), float(co.mean())]}}


def parity_check(args, rank, world, local, ctx, dist, torch, _capi):
    """Outside the timed region, for every N: a ~2e5-particle C4 box sharded over the N ranks like the bench
    workload, two steps, compared with the single-process oracle on rank 0.  Twice: in the configuration the
    bench times (fast fp, radix sort: every main-layout field and dt, relative error per particle) and in the
    bit-exact configuration (strict fp, the reference's bitonic tie order: Morton codes, permutation, tree,
    neighbour lists and every field identical)."""
    from tests import scenarios as S

    n_target = args.parity_npart
    out = {"n": None, "world": world, "steps": 2}

    def run(fp_mode, sort_mode):
        sc = S.periodic_box(n_target, "M4", "cd10", jitter=0.1, grid=(world, 1, 1), stretch=(world, 1, 1),
                            sort_mode=sort_mode)
        ids = [_capi.nccl_unique_id() if rank == 0 else None]
        if world > 1:
            dist.broadcast_object_list(ids, src=0)
        # the fast run is the configuration the bench times (keep_step_data off: adaptive list tolerance, §4c)
        m = S.make_cuda(sc, ctx=ctx, keep_step_data=(fp_mode == "strict"), rank=rank, world=world, nccl_id=ids[0],
                        fp_mode=fp_mode)
        for _ in range(2):
            sm = m.evolve_once()
        ltol = m.list_tolerance()["last"]
        names = CHECK_FIELDS + (INT_NAMES if fp_mode == "strict" else [])
        mine = {ip: {nm: m.get(ip, nm) for nm in names} for ip in range(m.patch_count)
                if m.patch_is_local(ip) and m.patch_size(ip)}
        m.close()
        parts = [None] * world
        if world > 1:
            dist.all_gather_object(parts, mine)
        else:
            parts = [mine]
        if rank != 0:
            return None
        got = {}
        for d in parts:
            got.update(d)
        oracle_all_cores()
        o = S.make_oracle(sc)
        for _ in range(2):
            so = o.evolve_once()
        worst, worst_nm, ints_ok, bits_ok = 0.0, "", True, True
        for ip in range(o.patch_count):
            if o.patch_size(ip) == 0:
                continue
            for nm in names:
                a, b = got[ip][nm], o.get(ip, nm)
                if nm in INT_NAMES:
                    ints_ok = ints_ok and a.shape == b.shape and bool(np.array_equal(a, b))
                    continue
                e = rel_err(a, b)
                bits_ok = bits_ok and bool(np.array_equal(a, b))
                if e > worst:
                    worst, worst_nm = e, f"{nm}@patch{ip}"
        e_dt = abs(sm["dt"] - so["dt"]) / abs(so["dt"])
        if e_dt > worst:
            worst, worst_nm = e_dt, "dt"
        return {"n": len(sc["xyz"]), "max_rel_err": worst, "worst": worst_nm, "ints_exact": ints_ok,
                "fields_bit_identical": bits_ok, "h_subcycles": sm["h_subcycles"] == so["h_subcycles"],
                "list_tolerance": ltol}

    fast = run(args.fp, "radix")
    strict = run("strict", "bitonic")
    if rank != 0:
        return None
    out.update(n=fast["n"], max_rel_err=fast["max_rel_err"], worst=fast["worst"],
               ints_exact=strict["ints_exact"], strict_fields_bit_identical=strict["fields_bit_identical"],
               strict_max_rel_err=strict["max_rel_err"], list_tolerance_second_step=fast["list_tolerance"],
               config={"bench_mode": f"fp {args.fp}, radix sort", "exact_mode": "fp strict, bitonic sort (reference tie order)",
                       "tolerance": "1e-10 relative per particle, |d| <= tol * max(|x|, mean|x|), no floors"})
    return out


_JSON_OUT = None


def claim_stdout():
    """stdout carries ONE JSON line: whatever libraries print there (NCCL's version banner, torchrun chatter of the
    children) is sent to stderr; the line itself goes to the original descriptor."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    claim_stdout()
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--npart-per-gpu", type=int, default=16 * 2**20)
    ap.add_argument("--cpu-sample", type=int, default=4 * 2**20,
                    help="particles of the CPU arm's sample box (BASELINE.md §3: a subset box, scaled linearly)")
    ap.add_argument("--parity-npart", type=int, default=200000)
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--fp", default="fast", choices=["fast", "strict"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--host-setup", action="store_true",
                    help="build the initial conditions with numpy on the host instead of on the device")
    ap.add_argument("--no-reorder", action="store_true",
                    help="keep the generation order of the lattice (the reference's apply_setup reorders by default)")
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from shamrock_b200 import _capi
    from tests import scenarios as S

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    cpus_before = bind_to_gpu_numa_node(local)
    nccl_id = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ids = [_capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        nccl_id = ids[0]

    def count_reduce(n_local):
        t = torch.tensor([n_local], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        return int(t.item())

    ctx = _capi.Context(local)
    t_setup = time.perf_counter()
    if args.host_setup:  # numpy lattice on the host, pushed patch by patch (round-1 path)
        sc = workload(args.npart_per_gpu, world, rank=rank, count_reduce=count_reduce if world > 1 else None)
        m = S.make_cuda(sc, ctx=ctx, keep_step_data=False, rank=rank, world=world, nccl_id=nccl_id, fp_mode=args.fp)
    else:  # generated on the device, every rank its own patches (shamb200_model_add_lattice_hcp + setters)
        sc = S.periodic_box(args.npart_per_gpu * world, "M4", "cd10", jitter=0.0, grid=(world, 1, 1),
                            stretch=(world, 1, 1), sort_mode="radix", local_boxes="device")
        m = S.make_cuda(sc, ctx=ctx, keep_step_data=False, rank=rank, world=world, nccl_id=nccl_id, fp_mode=args.fp)
        S.periodic_box_on_device(m, args.npart_per_gpu * world, "M4", stretch=(world, 1, 1))
    ctx.synchronize()
    t_setup = time.perf_counter() - t_setup
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    if not args.no_reorder:  # SPHSetup::apply_setup(part_reordering = true), the protocol's default
        m.reorder_particles()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    per_step_ms = []

    def timed(fn, k):
        """device time of k calls (events on the library's stream; max over ranks); the per-call times of the last
        measurement are left in per_step_ms (reference protocol: best and median of the replays)"""
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
        ev[0].record(stream)
        for i in range(k):
            fn()
            ev[i + 1].record(stream)
        ev[k].synchronize()
        barrier()
        ms = ev[0].elapsed_time(ev[k])
        per_step_ms[:] = [ev[i].elapsed_time(ev[i + 1]) for i in range(k)]
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def step():
        m.set_next_dt(0.0)
        m.evolve_once()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    fp64_peak = ctx.microbench("fp64")
    copy_bw = ctx.microbench("copy")
    # 32-byte record gathers from an L1-resident table (microbench.cu): the rate of the unit that binds the SPH
    # neighbour loops, for a random record per lane (what a neighbour list is) and for the conflict-free case
    gather_random, gather_aligned = ctx.microbench(11), ctx.microbench(12)

    # warm-up: first real timestep (converges h), then dt = 0 replays
    m.evolve_once()
    for _ in range(args.warmup - 1):
        step()
    st = m.state()
    n_local = int(st["n_local"])
    n_total = int(st["npart"])

    _capi.reset_launch_count()
    stage_acc = {}

    def step_acc():
        step()
        for k, v in m.stage_times().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v

    clocks.begin()
    ms = timed(step_acc, args.steps)
    clocks.end()
    step_times = sorted(per_step_ms)
    launches = _capi.launch_count()
    clk = clocks.stop() if rank == 0 else None
    st = m.state()
    value = n_total * args.steps / (ms * 1e-3)

    # ---- e2e: host patch data in, host patch data out, every step ---------------------------------
    e2e = None
    if not args.no_e2e:
        # the reference-facing call: shamb200_model_evolve_once_host on page-locked HOST patch data.  Inputs
        # of a step (xyz vxyz axyz hpart uint duint alpha_AV) go up, all 12 main-layout fields come back.
        host, h2d, d2h = {}, 0, 0
        ips = [ip for ip in range(m.patch_count) if m.patch_is_local(ip) and m.patch_size(ip)]
        IN = ["xyz", "vxyz", "axyz", "hpart", "uint", "duint", "alpha_AV", "soundspeed"]
        # every field the step writes comes back; axyz_ext is not one of them in this configuration (no external
        # force: the host's copy stays the zeros it holds), its out pointer is NULL
        OUT = [nm for nm, _ in MAIN_FIELDS if nm != "axyz_ext"]
        for ip in ips:
            for nm, nv in MAIN_FIELDS:
                a = m.get(ip, nm)
                t = torch.empty(a.size, dtype=torch.float64).pin_memory()
                t.numpy()[:] = a.reshape(-1)
                host[(ip, nm)] = t

        def e2e_step():
            nonlocal h2d, d2h
            m.set_next_dt(0.0)
            h2d = d2h = 0
            for ip in ips:
                m.evolve_once_host(ip, m.patch_size(ip), {nm: host[(ip, nm)].data_ptr() for nm in IN},
                                   {nm: host[(ip, nm)].data_ptr() for nm in OUT})
                a, b = m.host_traffic()
                h2d, d2h = h2d + a, d2h + b

        e2e_step()
        k2 = max(2, min(args.steps, 3))
        ms2 = timed(e2e_step, k2)
        e2e_val = n_total * k2 / (ms2 * 1e-3)
        tot = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tot)
        e2e = {"value": e2e_val, "unit": "particles/s", "h2d_bytes_per_step": int(tot[0].item()),
               "d2h_bytes_per_step": int(tot[1].item()), "ms_per_step": ms2 / k2,
               # bytes of a step over the step's wall time: what the host link sustains while the kernels run
               # (the copies overlap the kernels and each other: the two directions are concurrent)
               "h2d_GBps_over_step": float(tot[0].item()) / (ms2 / k2 * 1e-3) / 1e9,
               "d2h_GBps_over_step": float(tot[1].item()) / (ms2 / k2 * 1e-3) / 1e9,
               "api": "shamb200_model_evolve_once_host (pinned host patch data; copies on two copy streams, "
                      "overlapped with the kernels)",
               "fields_up": IN, "fields_down": OUT,
               # id ranges the operator / force / corrector passes of the host step are cut into (finished ranges
               # travel while the next is computed; needs Morton-ordered patch data: ParticleReordering)
               "id_ranges": m.host_step_info(ips[0])[0] if ips else 0}

    if rank == 0:
        hbm_peak, peak_src = peaks()
        # stage = device time between CUDA-event marks on the step's stream; each heavy stage is ONE kernel
        # (h_solve / av_operators / force_cfl) or the search kernels.  Roof = slower of HBM and FP64 pipe.
        N, K = n_local, int(st["K_local"])
        tests = m.search_stats()[1]
        ltol_info = m.list_tolerance()
        ltol = ltol_info["last"] or 1.1
        M, L = int(N * 1.0), int(N / 3.6)  # ~3.6 objects per leaf at reduction level 3 on the HCP lattice
        sweeps = int(st["h_iters_last"]) + 1
        per_stage = {k: v / args.steps for k, v in stage_acc.items()}
        table = {}
        for k, ms_k in per_stage.items():
            w = alg_work(k, N, M, K, L, sweeps, tests, list_tol=ltol)
            if not w:
                continue
            t_hbm, t_fp = w[0] / (hbm_peak * 1e9), (w[1] / (fp64_peak * 1e12) if fp64_peak else 0.0)
            bound = "fp64" if t_fp > t_hbm else "hbm"
            table[k] = {"ms": round(ms_k, 3), "bound": bound, "frac": max(t_hbm, t_fp) / (ms_k * 1e-3),
                        "GB/s": w[0] / (ms_k * 1e-3) / 1e9, "TFLOP/s": w[1] / (ms_k * 1e-3) / 1e12}
        # records gathered per launch by the three neighbour loops (one 32-byte record per list entry and array)
        for k, nrec in (("h_iteration", sweeps * K), ("divv_curlv_dtdivv", 3 * K), ("forces", 3 * K)):
            if k in table:
                rate = nrec / (table[k]["ms"] * 1e-3) / 1e9
                table[k]["gather"] = {"records": nrec, "Grec/s": rate, "frac_of_l1_random_gather": rate / gather_random}
        top = max(table, key=lambda k: table[k]["ms"])
        tt = table[top]
        # DRAM bytes of one launch of that kernel from the committed ncu --set full capture of this workload;
        # the same capture gives, per heavy kernel, the utilisation of the units that actually bind it (the
        # neighbour loops saturate the L1 data pipe of the LSU - 32-byte record gathers - before the FP64 pipe)
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic_17M.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if abs(tj["npart"] - n_local) <= 0.01 * n_local:
                def fresh(k):  # a capture of other code (kernel time off by > 5 %) is not evidence for this run
                    v = tj["stages"][k]
                    return k in table and abs(v["ncu_time_ms"] - table[k]["ms"]) <= 0.05 * table[k]["ms"]

                if top in tj["stages"] and fresh(top):
                    traffic = tj["stages"][top]["traffic"]
                    traffic_src = "profiles/traffic_17M.json (" + tj["stages"][top]["kernel"] + ")"
                for k, v in tj["stages"].items():
                    if k in table and fresh(k):
                        table[k]["ncu"] = {q: v[q] for q in ("fp64_pipe_pct", "l1_lsu_data_pipe_pct", "issue_active_pct",
                                                             "ncu_time_ms") if q in v}
                    elif k in table:
                        table[k]["ncu"] = {"stale": f"capture {v['ncu_time_ms']:.2f} ms vs {table[k]['ms']:.2f} ms live"}
        roofline = {"bound": tt["bound"], "kernel": top,
                    "achieved": tt["TFLOP/s"] if tt["bound"] == "fp64" else tt["GB/s"],
                    "peak": fp64_peak if tt["bound"] == "fp64" else hbm_peak,
                    "unit": "TFLOP/s" if tt["bound"] == "fp64" else "GB/s", "frac": tt["frac"], "traffic": traffic, "traffic_source": traffic_src,
                    "algorithmic_bytes": alg_work(top, N, M, K, L, sweeps, tests, list_tol=ltol)[0],
                    "algorithmic_flops": alg_work(top, N, M, K, L, sweeps, tests, list_tol=ltol)[1],
                    "accept_tests_per_particle": tests / max(N, 1),
                    "neighbours_per_particle": K / max(N, 1),
                    # the lists of the timed steps were built with the radius R h list_tolerance.last instead of the
                    # reference's R h 1.1: the h growth of the previous step (dt = 0 replays: none) decides, a step
                    # whose h outgrows its lists is redone with 1.1 (shamb200_model_list_tolerance)
                    "list_tolerance": ltol_info,
                    "peak_source": {"hbm_gbs": hbm_peak, "hbm": peak_src, "fp64_tflops": fp64_peak,
                                    "fp64": "measured here: FP64 FMA chains (shamb200_microbench), FMA = 2 flop",
                                    "copy_gbs_here": copy_bw,
                                    "l1_gather_random_Grec_s": gather_random,
                                    "l1_gather_bank_aligned_Grec_s": gather_aligned,
                                    "l1_gather": "measured here: one 256-bit load per lane from a 32 KB table "
                                                 "(shamb200_microbench 11 / 12), G records of 32 B per second"},
                    "note": "roof = slower of the FP64 pipe and HBM (north star); flops: FMA=2, div/sqrt=1; "
                            "stages[*].ncu = unit utilisation from the committed ncu capture (profiles/traffic_17M.json): "
                            "the neighbour loops run at 75-97 % of the L1 LSU data pipe",
                    "ms_per_launch": tt["ms"], "stages": table,
                    "stage_ms": {k: round(v, 3) for k, v in per_stage.items()}}
        line = {
            "metric": "SPH particle-updates/sec (full step)", "value": value, "unit": "particles/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "ms_per_step_best": step_times[0], "ms_per_step_median": step_times[len(step_times) // 2],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": value / H100_PUBLISHED, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "C4: periodic HCP box, M4 kernel, CD10 AV, adiabatic gamma=5/3, Sedov-like uint "
                       "injection, dt=0 replay (reference protocol sph_homogeneous_benchmark.py)",
                       "npart_total": n_total, "npart_per_gpu": n_total // world, "neighbours_per_particle": K / max(N, 1),
                       "patches": list(sc["grid"]), "sort": sc["sort_mode"],
                       "ic": ("host numpy lattice" if args.host_setup else
                              "generated on the device (shamb200_model_add_lattice_hcp, set_value_in_a_box, "
                              "add_kernel_value)") + f", {t_setup:.2f} s",
                       "setup": "lattice order" if args.no_reorder else
                       "patch data Morton-reordered once after the setup (apply_setup part_reordering=True, the "
                       "reference's default)",
                       "fp": {"fast": "fast (FMA, per-particle reciprocals; parity 1e-10 relative vs the oracle)",
                              "strict": "strict (no FMA, bit-identical to the oracle)"}[args.fp],
                       "l2": "inputs larger than L2 (no flush needed)",
                       "vs_baseline_ref": "reference on 1x H100, 25.5 M part/s (BASELINE.md §1)"},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "h_subcycles": st["h_subcycles"], "corrector_iter": st["corrector_iter"],
        }
    m.close()
    pc = None
    if not args.no_parity_check:
        if rank == 0 and cpus_before:
            os.sched_setaffinity(0, cpus_before)  # the oracle uses every host core again
        pc = parity_check(args, rank, world, local, ctx, dist, torch, _capi)
    if rank == 0:
        line["parity_check"] = pc
        if not args.no_cpu_baseline and world == 1:
            if cpus_before:
                os.sched_setaffinity(0, cpus_before)
            line["cpu_baseline"] = cpu_baseline(args, ctx)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
