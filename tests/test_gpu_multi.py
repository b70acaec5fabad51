"""Row (e): ratings sharded by nonzero over >= 2 GPUs, NCCL row exchange; result equals the oracle and
every rank ends with bit-identical replicated factors.  Skipped on single-GPU boxes."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("exchange", ["p2p", "nccl", "mc"])
@pytest.mark.parametrize("world", [2])
def test_sharded_fit_matches_oracle(world, exchange):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(HERE, "multi_gpu_worker.py"), exchange]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("MULTI_GPU_OK") == world
