"""CPU tests of the multi-GPU host layer's control block (csrc/dist.cu): POSIX shared-memory rendezvous, barriers,
all-gathers, splitter choice and staging-layout arithmetic across real processes -- no CUDA call involved."""
import ctypes as C
import importlib
import multiprocessing as mp
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(session, rank, world, q):
    sys.path.insert(0, ROOT)
    dbt = importlib.import_module("database-technology-algorithms_b200")
    chk = C.c_uint64()
    rc = dbt.lib().dbt_dist_selftest_host(session.encode(), rank, world, C.byref(chk))
    q.put((rank, rc, chk.value))


@pytest.mark.parametrize("world", [1, 2, 3, 4])
def test_control_block_rendezvous_barriers_allgathers_splitters_layout(dbt, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    session = f"selftest_{os.getpid()}_{world}"
    procs = [ctx.Process(target=_worker, args=(session, r, world, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r for r, _, _ in res] == list(range(world))
    assert all(rc == 0 for _, rc, _ in res), res
    assert len({chk for _, _, chk in res}) == 1, res  # every rank saw the same messages, splitters and layouts
    assert not os.path.exists(f"/dev/shm/dbt_{session}")  # the control block's name is removed once everybody is attached


def test_a_missing_rank_is_a_timeout_not_a_hang(dbt):
    # rank 1 of a 2-rank group whose rank 0 never shows up: the rendezvous gives up (30 s bound inside the self test)
    chk = C.c_uint64()
    rc = dbt.lib().dbt_dist_selftest_host(f"nobody_{os.getpid()}".encode(), 1, 2, C.byref(chk))
    assert rc == -7  # DBT_ERR_TIMEOUT
