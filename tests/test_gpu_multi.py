"""Multi-GPU parity: the time-sharded path over 2 GPUs (NCCL) against the single-GPU golden results."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("exchange", ["default", "nccl"])
def test_two_gpu_sharded_matches_golden(exchange):
    """exchange: the A12 sub-strips reach their owner rank through peer memory (the map-side kernel's own NVLink
    stores; the default where CUDA IPC mapping works) or through the ncclSend/ncclRecv all-to-all"""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_check.py")]
    env = dict(os.environ)
    if exchange == "nccl":
        env["EMBA_XCHG_PEER"] = "0"
        env["EMBA_EXPECT_EXCHANGE"] = "nccl"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    assert "mgpu_check small: world=2 OK" in r.stdout
