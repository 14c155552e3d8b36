"""Data-parallel TrainStep on real GPUs (needs >= 2 devices; skipped on a one-GPU box): launches tools/ddp_check.py under
torchrun -- replicas start from rank 0's weights, end every step bit-identical, and equal the averaged-gradient update."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_train_step_matches_averaged_gradient_update():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "ddp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DDP CHECK OK" in r.stdout, r.stdout[-4000:] + r.stderr[-2000:]
