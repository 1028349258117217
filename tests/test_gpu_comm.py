"""kvm_comm_* / kvm_gather_result: the multi-GPU tail inside the library, one process per GPU (needs two GPUs: NCCL
refuses two ranks on one device)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("p2p", ["1", "0"])
def test_library_gather_two_ranks(p2p):
    """p2p = 1: the fixed-size round runs as the peer-memory exchange kernel (CUDA IPC), the overflow round over NCCL;
    p2p = 0: both rounds over NCCL."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533" if p2p == "1" else "29534", os.path.join(ROOT, "tests", "_comm_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, KVM_GATHER_P2P=p2p))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MISMATCH" not in r.stdout and r.stdout.count("OK") == 4, r.stdout
