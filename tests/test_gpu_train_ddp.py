"""Data-parallel training step on >= 2 GPUs (NCCL): tools/ddp_check.py under torchrun.  Skipped on a single-GPU box."""
import pathlib
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_reduced_gradients_equal_full_batch_and_ranks_stay_in_sync():
    n = min(torch.cuda.device_count(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29761", str(ROOT / "tools" / "ddp_check.py")],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "DDP_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
