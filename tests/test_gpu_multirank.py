"""Multi-GPU parity (needs >= 2 GPUs; run with `gpurun --gpus 2`): patches sharded over two ranks, ghost
zones / migration over NCCL send/recv, dt / eps_v / convergence over NCCL allreduce; every rank's patches
must equal the single-process oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("fp_mode", ["strict", "fast"])
@pytest.mark.parametrize("which", ["periodic", "sod", "disc", "disc_balanced", "scheduler"])
def test_two_rank_step_matches_oracle(which, fp_mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "mp_step_worker.py"), which, fp_mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-6000:]
    assert r.stdout.count(f"{which} {fp_mode} ok") == 2


@pytest.mark.parametrize("which", ["adaptive", "adaptive_fallback"])
def test_two_rank_adaptive_list_tolerance(which):
    """fast fp mode with keep_step_data off (what bench.py times): the list tolerance follows the h growth and is
    the same on both ranks; main-layout fields against the oracle at 1e-10 over five real steps"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29518", os.path.join(ROOT, "tests", "mp_step_worker.py"), which, "fast"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-6000:]
    assert r.stdout.count(f"{which} fast ok") == 2
