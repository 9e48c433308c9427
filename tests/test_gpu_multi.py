"""GPU tier, needs >= 2 GPUs (skipped otherwise): the NCCL row-sharded path equals the single-GPU sweep."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_sweep_equals_single_gpu_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "sharded == whole: True" in r.stdout and "dist_check: ok" in r.stdout
    assert "ok=False" not in r.stdout and ": False" not in r.stdout
