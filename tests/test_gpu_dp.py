"""Data-parallel parity on real GPUs (needs >= 2 devices; skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_dp2_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(HERE, "dp_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "DP_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
