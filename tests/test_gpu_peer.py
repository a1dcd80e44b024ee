"""Multi-GPU: the gather fused into the rollout (peer stores over NVLink, fancy_gym_b200.dist.PeerResultExchange) against an
NCCL all-gather of the same blocks.  Needs two GPUs on the box (skipped otherwise); world-size-2 logic of the NCCL path is
covered on CPU with gloo (tests/test_dist_gloo.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_store_gather_equals_nccl_all_gather():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "check_peer_exchange.py")],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and "PEER OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
