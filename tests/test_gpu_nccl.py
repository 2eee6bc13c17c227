"""The N > 1 path on real GPUs: two ranks under torchrun, NCCL backend, pooled posterior VALUES checked on every
rank (tests/nccl_pool_worker.py).  Needs two visible GPUs (`gpurun --gpus 2`); the gloo twin runs on CPU
(tests/test_chains.py)."""
import os
import socket
import subprocess
import sys

import pytest

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
@pytest.mark.timeout(600)
def test_pool_posterior_values_over_nccl():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "nccl_pool_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=550, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "NCCL_POOL_OK world=2" in res.stdout
