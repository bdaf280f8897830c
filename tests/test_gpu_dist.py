"""-m gpu, needs >= 2 GPUs (skipped otherwise): the CUDA multi-GPU operators under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("exchange", ["p2p", "nccl", "p2p-records", "p2p-keys"])
def test_two_gpu_operators_match_the_oracle(exchange):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, DBT_DIST_EXCHANGE=exchange.split("-")[0], DBT_DIST_SORT={"records": "records", "keys": "keys"}.get(exchange.split("-")[-1], "overlap"))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", {"p2p": "29533", "nccl": "29534", "p2p-records": "29535", "p2p-keys": "29536"}[exchange], os.path.join(HERE, "dist_check.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert "DIST_CHECK_PASSED" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
