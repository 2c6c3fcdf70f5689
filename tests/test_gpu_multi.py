"""Multi-GPU parity (needs >= 2 visible GPUs, skipped otherwise): the sharded plan (R = 2) is bit-identical to the
unsharded plan (R = 1, same seed) on the CUDA path, and the sharded plan with the cost exchange fused into the
cost kernel over NVLink peer memory is bit-identical to the NCCL all-gather path and identical on every rank. Spawns
tests/gpu_peer_gather_check.py under torchrun (one process per GPU, 127.0.0.1 rendezvous)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_memory_cost_exchange_matches_nccl():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(root, "tests", "gpu_peer_gather_check.py")]
    res = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "peer-memory cost exchange == NCCL all-gather" in res.stdout
    assert "sharded plan (R = 2) == unsharded plan (R = 1)" in res.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_data_parallel_training_step():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29543", os.path.join(root, "tests", "gpu_dp_train_check.py")]
    res = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "overlapped all-reduce == one all-reduce" in res.stdout
