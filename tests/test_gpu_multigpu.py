"""Runs the sharded-cluster check under torchrun when the box has at least two GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_house_sharded_cluster_over_two_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(here, "run_sharded_multigpu.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "sharded multi-GPU check OK" in res.stdout
