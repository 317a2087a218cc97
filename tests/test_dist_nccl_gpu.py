"""(e) on real GPUs: sharded vs single-GPU parity under torchrun + NCCL.  Needs >= 2 GPUs (gpurun --gpus 2)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["p2p", "a2a"])
@pytest.mark.parametrize("k", [16, 64])
def test_sharded_deepfm_matches_single_gpu(k, mode):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ, DIST_K=str(k), DIST_MODE=mode)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + k + (1 if mode == "p2p" else 0)), os.path.join(ROOT, "scripts", "dist_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "dist_check ok" in r.stdout
