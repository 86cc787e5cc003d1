"""Data-parallel parity on real GPUs (needs >= 2 devices; skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("comm", ["peer", "nccl"])
def test_dp2_equals_single_gpu(comm):
    """2-GPU data-parallel step (NVLink peer-fused exchange+Adam, or NCCL all-reduce + local Adam) == the
    single-GPU step on the same global batch: loss, reduced gradients and updated parameters <= 1e-5."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29611" if comm == "peer" else "29612", os.path.join(HERE, "dp_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, SN_DP_COMM=comm))
    assert "DP_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_dp2_early_vocab_exchange_equals_single_exchange():
    """bf16 mode, peer-fused exchange: the vocabulary projection's reduce-scatter + Adam + all-gather issued early (under
    the reverse recurrence) == one exchange at the end of the step; all ranks end with identical parameters."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29613", os.path.join(HERE, "dp_early_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert "DP_EARLY_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
