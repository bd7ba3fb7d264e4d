"""Tile-sharded path on real GPUs (needs >= 2): one process per GPU under torchrun, NCCL
all-gather of the resampled rows and of the edge counts; union of the shards == oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_sharded_pipeline():
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if ng < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29611",
                        os.path.join(ROOT, "tools", "multigpu_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("multigpu_check ok") == 4   # DMMA, FMA, tcgen05 one-shot, tcgen05 streamed
