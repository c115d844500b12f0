"""GPU tests that need MORE THAN ONE device (skipped on a one-GPU box; the data-parallel host logic is covered on CPU by the
world_size-2 gloo tests in test_host_logic.py, the peer all-reduce protocol on one device by test_gpu_aux.py)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_data_parallel_peer_allreduce_two_processes():
    """One process per GPU (torchrun, NCCL for the set-up): the train step's exchange runs through qfa_peer_allreduce over
    symmetric memory -- bit-identical sums on all ranks, equal to the rank-ordered sum, round-off away from ncclAllReduce; the
    captured step graph keeps the replicas bit-identical.  See tests/workers/dp_peer_worker.py."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "workers", "dp_peer_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "DP-PEER-OK" in r.stdout, r.stdout[-3000:] + r.stderr[-6000:]
