"""Multi-GPU paths under test: data-parallel training (BASELINE config 3) and the sharded NLL / ELBO sweep (config 5)
on 2 GPUs of one node, launched exactly as the driver launches bench.py (torchrun, one process per GPU, NCCL over
NVLink).  Skipped on a single-GPU box; the CPU (gloo, world size 2) tests of the host logic are in
tests/test_cpu_distributed.py."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_data_parallel_step_matches_the_single_process_global_batch():
    """tests/multi/dp_worker.py: xrank_sum bit-identical on all ranks; sharded step == single-process step at the
    global batch (loss <= 1e-4, BatchNorm running statistics <= 1e-4, near-loss gradients), NCCL and peer-memory paths;
    ranks draw distinct noise / timesteps and keep identical parameters; sharded NLL + ELBO == single process."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi", "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=420)
    assert r.returncode == 0 and "DP-WORKER OK world=2" in r.stdout, (r.stdout[-3000:] + r.stderr[-3000:])
