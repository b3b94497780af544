"""N > 1 on real GPUs: launches tests/multi_gpu_check.py under torchrun when >= 2 GPUs are visible."""
import os
import subprocess
import sys

import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def test_two_rank_parity_over_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29531",
                        os.path.join(H.ROOT, "tests", "multi_gpu_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multi-GPU parity ok" in r.stdout
