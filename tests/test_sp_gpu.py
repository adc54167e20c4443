"""GPU (needs >= 2 devices, skipped otherwise): Ulysses sequence parallelism over NCCL, forward + backward, against
the oracle — runs tests/sp_check.py under torchrun."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ulysses_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29655", os.path.join(ROOT, "tests", "sp_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(res.stdout[-3000:], res.stderr[-2000:])
    assert res.returncode == 0 and "SP CHECK OK" in res.stdout
