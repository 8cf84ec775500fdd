"""Row-block sharded head, world_size 2: NCCL when >= 2 GPUs are visible, otherwise both ranks share GPU 0 with
gloo-staged collectives (no device-side waiting between ranks, so co-scheduling on one GPU is safe)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sharded_head_world2(precision):
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dist_sharded_check.py"), "--backend",
           backend, "--precision", precision]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" OK ") == 2, r.stdout[-2000:]


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_graph_captured_sharded_step_world2(precision):
    """GraphedHeadStep at world_size 2 over NCCL (captured collectives) against the single-process full-batch head,
    and a later replay against the eager sharded step.  Needs one GPU per rank."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (NCCL collectives captured in a CUDA graph; gloo cannot be captured)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "tests", "dist_graph_check.py"), "--precision",
           precision]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" OK ") == 2, r.stdout[-2000:]
