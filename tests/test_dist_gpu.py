"""Multi-GPU parity (needs >= 2 visible GPUs; skipped otherwise): runs tests/dist_check.py under torchrun with NCCL:
K-/N-sharded GEMVs vs the single-GPU result, the fused push all-reduce chain vs the NCCL chain, graph == eager, and
the tensor-parallel full decode step vs the single-GPU model."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_paths_world2():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-u", "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29733", os.path.join(ROOT, "tests", "dist_check.py"), "nccl"]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=400, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "DIST CHECK OK" in res.stdout


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 4, reason="needs 4 GPUs")
def test_sharded_paths_world4():
    """world size 4: the persistent engine is the default here; K cuts fall inside 1024-chunks (re-packed shards)"""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-u", "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr",
           "127.0.0.1", "--master-port", "29734", os.path.join(ROOT, "tests", "dist_check.py"), "nccl"]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=500, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "DIST CHECK OK" in res.stdout
