"""GPU, >= 2 devices: fused render+gather over peer memory vs render + NCCL all_gather (launched through torchrun)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_peer_gather_matches_nccl_all_gather():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, "tools", "check_peer_gather.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "PEER_GATHER_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
