"""Multi-process, multi-GPU parity (needs >= 2 devices; the single-GPU box of the round-end run skips it): the
library's shard group (silo_gpu_shard_group_*) across real devices and processes, against the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_group_across_processes():
    import torch
    n_gpus = torch.cuda.device_count()
    if n_gpus < 2:
        pytest.skip("needs at least two GPUs")
    world = min(n_gpus, 4)
    result = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
         "--master-port", "29517", os.path.join(ROOT, "tests", "multi_gpu_shard_group.py")],
        capture_output=True, text=True, timeout=600)
    assert result.returncode == 0, result.stdout[-3000:] + result.stderr[-3000:]
    assert "multi-GPU shard group ok" in result.stdout
