"""Multi-GPU (NCCL, one process per GPU) checks; skipped unless the box has at least two GPUs.  The same checks
run stand-alone under torchrun as tests/multi_gpu_check.py (validated at 2 and 8 B200s)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_nccl_check_two_ranks():
    """fused step over the packed (peer-memory) all-gather == oracle; fused pack+gather == pack_pair + NCCL all-gather
    bit for bit; column- and row-sharded retrieval == single GPU; negative-row exchange == gather-with-grad + index."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=500, cwd=ROOT)
    assert r.returncode == 0 and "nccl_check OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
