"""Row (e): ratings sharded by user range over 2 / 4 / 8 GPUs; in-switch ("mc"), copy-engine staged ("ce") and NCCL
combine of the item pass's row sums.  Results equal the oracle and every rank ends with bit-identical replicated tables.  Each world size is
skipped on boxes with fewer GPUs."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("exchange", ["mc", "ce", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_fit_matches_oracle(world, exchange):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs >= {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29533 + world), os.path.join(HERE, "multi_gpu_worker.py"), exchange]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("MULTI_GPU_OK") == world
