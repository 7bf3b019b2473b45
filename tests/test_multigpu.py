"""GPU (>= 2 devices): the collective build path (NCCL id exchange, VMM shard mapping with the
legacy CUDA-IPC fallback), in-kernel peer loads and the NCCL id-exchange extract, on the reference's 2-rank scenarios and a sharded parity check."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_rank_worker(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mp_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out[-4000:]
    for r in range(world):
        assert f"RANK {r} OK" in out, out[-4000:]
