"""Sharded paths over NCCL on real GPUs: scripts/nccl_check.py under torchrun, one rank per GPU.
Needs >= 2 CUDA devices (`gpurun --gpus 2`); skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two CUDA devices")
def test_sharded_attack_and_header_over_nccl():
    """sharded_attack (emb, e2e; batch sliced over the ranks, global MSE normaliser, all_gather + all_reduce after the
    loop), its pre-sliced form and sharded_header_optimize (per-iteration all_reduce of the header gradient) equal the
    unsharded calls on one GPU; the data-parallel VSMask trainer (BatchNorm over the global batch, gradient all-reduce through
    the library's callback) equals one GPU on the concatenated batch."""
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "scripts", "nccl_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "FAIL" not in r.stdout and r.stdout.count("PASS") >= 4
