"""N > 1 on real GPUs (NCCL): needs >= 2 visible B200s (`gpurun --gpus 2`), skipped otherwise."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_matches_single_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tests", "dist_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "DIST_OK 2" in res.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_full_size_steps_match_the_cpu_oracle():
    """cfg5's parity variant (3 classes, 192x160x160, one instance per GPU) and cfg2's per-GPU batch, data-parallel
    over 2 ranks, against oracle.dp_loss.run_loss on the gathered global batch (see dist_worker.full_size_case)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29532", os.path.join(ROOT, "tests", "dist_worker.py"), "cfg5", "cfg2"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "DIST_OK 2" in res.stdout and "FAIL" not in res.stdout
