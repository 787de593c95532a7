"""Multi-GPU parity as a `-m gpu` test: launches tests/multi_gpu_check.py with one process per GPU (torchrun, NCCL)
on two devices and requires it to pass -- sharded closed-loop encode bit-identical to one GPU and to the oracle,
k-means ranks bit-identical to each other and within 1e-11 of the oracle.  Skipped on a single-GPU box (the gloo
world_size-2 twin of the exchange step runs in tests/test_host.py on the CPU)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_shards_bit_identical():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (this box has %d)" % torch.cuda.device_count())
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "bit-identical to 1 GPU" in p.stdout and "vq_train: ranks agree" in p.stdout
