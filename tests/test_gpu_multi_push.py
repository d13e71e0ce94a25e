"""Two ranks, two GPUs: fused pack + all-gather over NVLink peer memory against the NCCL all-gather
(SURVEY.md 8(e)).  Skipped on a single-GPU box; the host-side sharding logic is covered on CPU by
tests/test_multigpu_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_push_gather_equals_nccl_all_gather():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", os.path.join(root, "tests", "mgpu_push_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root)
    assert "PUSH_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
