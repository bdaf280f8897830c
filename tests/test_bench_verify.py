"""CPU tests of bench_verify.py (the independent torch checkers bench.py runs at full scale): they must accept the
oracle's results and reject corrupted ones."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_verify as V  # noqa: E402


def t(blocks):
    return torch.from_numpy(np.ascontiguousarray(blocks).view(np.uint8).reshape(-1).copy())


@pytest.fixture(scope="module")
def data(orc):
    r = orc.gen_syn(3, 20000, 6000, 1)
    s = orc.gen_syn(4, 30000, 6000, 1)
    return r, s


def test_columns_and_hashes_follow_the_record_layout(orc, data):
    r, _ = data
    rows = orc.rows_of(r)
    assert np.array_equal(V.column(t(r), len(rows), 0).numpy(), rows["recid"].astype(np.int64))
    assert np.array_equal(V.column(t(r), len(rows), 1).numpy(), rows["num"].astype(np.int64))
    h = V.row_hashes(t(r), len(rows), chunk_blocks=7).numpy()
    r2 = r.copy()
    e = r2["entries"].reshape(-1).copy()
    e["dummy2"][123] ^= 1                      # one bit in the last word of one record
    r2["entries"][:] = e.reshape(r2["entries"].shape)
    h2 = V.row_hashes(t(r2), len(rows)).numpy()
    assert (h != h2).sum() == 1 and h[123] != h2[123]
    k = V.str_key64(t(r), len(rows)).numpy()
    order = np.argsort(k, kind="stable")
    strs = [bytes(x).split(b"\0")[0] for x in rows["str"]]
    assert [strs[i] for i in order] == sorted(strs)


def test_sort_and_dedup_checks_accept_the_oracle_and_reject_corruption(orc, data):
    r, _ = data
    n = orc.count_rows(r)
    out = orc.sort(r, "1")
    res = V.check_sort(t(r), n, t(out), n, lambda img, m: V.column(img, m, 1))
    assert res["ordered_in_rank"] and res["record_multiset_hash_equal"] and res["rows"] == n
    bad = out.copy()
    e = bad["entries"].reshape(-1).copy()
    e[[10, 11]] = e[[11, 10]]
    bad["entries"][:] = e.reshape(bad["entries"].shape)
    res = V.check_sort(t(r), n, t(bad), n, lambda img, m: V.column(img, m, 1))
    assert not res["ordered_in_rank"] and res["record_multiset_hash_equal"]
    out2 = orc.sort(r, "2")
    assert V.check_sort(t(r), n, t(out2), n, V.str_key64)["ordered_in_rank"]
    d = orc.dedup(r, "1")
    u = orc.count_rows(d)
    res = V.check_dedup_u32(t(r), n, t(d), u, 1, key_space=1 << 16)
    assert all(res[k] for k in ("keys_strictly_ascending", "every_row_is_min_recid_of_its_key", "record_multiset_hash_equal")) and res["rows"] == u
    bad = d.copy()
    e = bad["entries"].reshape(-1).copy()
    dup = np.flatnonzero(np.bincount(orc.rows_of(r)["num"])[e["num"][:u]] > 1)[0]
    other = [x for x in orc.rows_of(r) if x["num"] == e["num"][dup] and x["recid"] != e["recid"][dup]][0]
    e[dup] = other                                # a row of the right key, but not the min-recid one
    bad["entries"][:] = e.reshape(bad["entries"].shape)
    res = V.check_dedup_u32(t(r), n, t(bad), u, 1, key_space=1 << 16)
    assert not res["every_row_is_min_recid_of_its_key"] and not res["record_multiset_hash_equal"]


def test_join_checks(orc, data):
    r, s = data
    nr, ns = orc.count_rows(r), orc.count_rows(s)
    out = orc.hashjoin(r, s, "1")
    k = orc.count_rows(out)
    assert V.check_semijoin_u32(t(r), nr, t(s), ns, t(out), k, 1, 6000)["same_rows_in_s_order"]
    assert V.check_semijoin_u32(t(r), nr, t(s), ns, t(out), k, 1, 6000, chunk_rows=4700)["same_rows_in_s_order"]  # S walked in chunks
    bad = out.copy()
    e = bad["entries"].reshape(-1).copy()
    e[[0, 1]] = e[[1, 0]]
    bad["entries"][:] = e.reshape(bad["entries"].shape)
    assert not V.check_semijoin_u32(t(r), nr, t(s), ns, t(bad), k, 1, 6000)["same_rows_in_s_order"]
    r3 = orc.gen_syn(21, 20000, 3000, 1)
    s3 = orc.gen_syn(21 ^ 0x5EED, 20000, 3000, 3)
    mj = orc.mergejoin(r3, s3, "3")[0]
    m = orc.count_rows(mj)
    res = V.check_mergejoin_composite(t(r3), 20000, t(s3), 20000, t(mj), m)
    assert m > 1000 and res["same_keys_and_recids_in_global_order"] and res["record_multiset_hash_equal"] and res["rows_expected"] == m
    res = V.check_mergejoin_composite(t(r3), 20000, t(s3), 20000, t(mj), m - 1)
    assert not res["same_keys_and_recids_in_global_order"]
