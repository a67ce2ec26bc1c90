"""Multi-GPU path on real GPUs (skipped with fewer than two): results written straight into rank 0's memory over
NVLink (parallel.PeerGather) are bit-identical to a single-GPU solve of the whole batch."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_peer_gather_matches_single_gpu_bitwise():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tests", "_peer_gather_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "peer gather ok 2" in res.stdout
