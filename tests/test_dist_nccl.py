"""The N > 1 path on real GPUs: one process per GPU over NCCL (world size 2), sharded bulk
scoring + the single all-reduce of the totals.  Needs two GPUs (`gpurun --gpus 2`); the host
logic of the same path is covered on CPU by tests/test_dist_gloo.py."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
def test_two_rank_nccl_bulk_scoring():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    assert "rank 0/2: ok" in out and "rank 1/2: ok" in out, out[-2000:]
