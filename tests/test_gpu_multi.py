"""Multi-rank parity: band-sharded and atmosphere-sharded results equal the single-GPU results bit for bit.
With >= 2 GPUs the ranks run one per GPU over NCCL; on a one-GPU box two ranks share cuda:0 and exchange over gloo,
so the sharded computation is verified wherever the GPU suite runs (the NCCL exchange itself is covered by
bench.py --gpus N, which reports "multi_gpu_bit_identical")."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3])
def test_ranks_bit_identical(world):
    import torch
    assert torch.cuda.is_available()
    if world > 2 and 2 <= torch.cuda.device_count() < world:
        pytest.skip("3 ranks need 3 GPUs or a single shared one")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(29517 + world), os.path.join(here, "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "banded_bit_identical=1 atmospheres_bit_identical=1" in r.stdout
