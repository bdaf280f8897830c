"""-m gpu, needs >= 2 GPUs (skipped otherwise): the C++ multi-GPU operators (dbt_dist_*) under torchrun, all four
operators x all four fields, block-dense and ragged shards, 1 and 3 key sub-ranges per owner, against the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("nproc", [2, 4])
def test_multi_gpu_operators_match_the_oracle(nproc):
    import torch

    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
                        "127.0.0.1", "--master-port", str(29533 + nproc), os.path.join(HERE, "dist_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert "DIST_CHECK_PASSED" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


@pytest.mark.parametrize("ndev", [2, 4])
def test_file_entry_points_on_several_gpus_write_the_oracle_files(ndev):
    """DBT_DEVICES=<n>: MergeSort / EliminateDuplicates / MergeJoin / HashJoin called exactly as main.cpp would, one
    process, n GPUs (one host thread each); the files must be byte-identical to the single-node oracle's."""
    import torch

    if torch.cuda.device_count() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    p = subprocess.run([sys.executable, os.path.join(HERE, "multi_file_check.py")], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, DBT_DEVICES=str(ndev)))
    assert "MULTI_FILE_CHECK_PASSED" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
