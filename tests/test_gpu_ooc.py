"""-m gpu tests of the out-of-core forms: the chunk size is forced far below the image
size so that small images take the runs + global-order + chunked-gather path (sort, dedup) and the
streamed-S path (hash join).  Bit-exact against the oracle, like the in-core operators."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H
from test_gpu_entrypoints import call_dedup, call_join, call_mergesort, read_blocks

pytestmark = pytest.mark.gpu
FIELDS = ["0", "1", "2", "3"]


@pytest.fixture()
def chunked(dbt):
    """Force out-of-core chunks of `blocks` blocks for the duration of a test."""
    L = dbt.lib()

    def set_chunk(blocks):
        dbt.check(L.dbt_host_set_chunk_blocks(blocks))

    yield set_chunk
    L.dbt_host_set_chunk_blocks(0)
    L.dbt_host_trim()


def ooc_stats(dbt):
    st = (C.c_uint64 * 6)()
    dbt.check(dbt.lib().dbt_host_ooc_stats(st))
    return dict(zip(("runs", "chunks", "shrinks", "widened", "staged", "s_chunks"), [int(x) for x in st]))


def host_sort(dbt, orc, blocks, field):
    out = orc.new_blocks(len(blocks))
    n = C.c_uint64()
    dbt.check(dbt.lib().dbt_host_mergesort(blocks.ctypes.data, len(blocks), ord(field), out.ctypes.data, 0, C.byref(n)))
    return out[: H.nb(n.value)], n.value


def host_dedup(dbt, orc, blocks, field):
    out = orc.new_blocks(len(blocks))
    n, u = C.c_uint64(), C.c_uint64()
    dbt.check(dbt.lib().dbt_host_dedup(blocks.ctypes.data, len(blocks), ord(field), out.ctypes.data, 0, C.byref(n), C.byref(u)))
    return out[: H.nb(u.value)], n.value, u.value


def host_hashjoin(dbt, orc, r, s, field, cap_blocks=None):
    cap = len(s) if cap_blocks is None else cap_blocks
    out = orc.new_blocks(max(cap, 1))
    n = C.c_uint64()
    rc = dbt.lib().dbt_host_hashjoin(r.ctypes.data, len(r), s.ctypes.data, len(s), ord(field), out.ctypes.data, cap, 0, C.byref(n))
    return rc, out[: H.nb(n.value)] if rc == 0 else None, n.value


@pytest.mark.parametrize("field", FIELDS)
def test_out_of_core_sort_and_dedup_match_the_oracle(dbt, orc, chunked, field):
    f1 = orc.gen_ref(31, 300, two=False, num_mod=5000)  # ~6 rows per num: runs share most of their keys
    rows = f1["entries"].reshape(-1)
    rng = np.random.default_rng(3)
    rows = rows[rng.permutation(len(rows))]              # recids are not in file order: the recid word is sorted too
    f1["entries"][:] = rows.reshape(f1["entries"].shape)
    chunked(64)                                           # 5 runs of <= 64 blocks
    got, n = host_sort(dbt, orc, f1, field)
    want = orc.sort(f1, field)
    assert n == 30000 and H.same_image(got, want), H.first_diff(got, want)
    st = ooc_stats(dbt)
    assert st["runs"] == 5 and st["chunks"] == 5 and st["shrinks"] == 0 and 300 <= st["staged"] <= 300 + 2 * 5 * 5
    got, n, u = host_dedup(dbt, orc, f1, field)
    want = orc.dedup(f1, field)
    assert (n, u) == (30000, orc.count_rows(want)) and H.same_image(got, want), H.first_diff(got, want)
    st = ooc_stats(dbt)
    assert st["runs"] == 5
    if field == "1":  # 5000 distinct nums spread over all five runs: one chunk's slices overflow the staging
        assert st["shrinks"] > 0 and st["chunks"] > 1


def test_out_of_core_with_ragged_blocks_few_keys_and_tiny_chunks(dbt, orc, chunked):
    f1 = orc.gen_ref(5, 257, two=False, num_mod=40)      # 40 distinct nums: every output chunk draws on every run
    f1["nreserved"][::7] = 13                             # ragged: some blocks hold 13 live rows
    f1["entries"]["valid"][::7, 13:] = 0
    f1["nreserved"][-1] = 1
    f1["entries"]["valid"][-1, 1:] = 0
    chunked(40)                                           # 7 runs
    for field in FIELDS:
        got, n = host_sort(dbt, orc, f1, field)
        want = orc.sort(f1, field)
        assert n == orc.count_rows(f1) and H.same_image(got, want), (field, H.first_diff(got, want))
        got, n, u = host_dedup(dbt, orc, f1, field)
        want = orc.dedup(f1, field)
        assert u == orc.count_rows(want) and H.same_image(got, want), (field, H.first_diff(got, want))
    # an empty image and an image of empty blocks
    empty = orc.new_blocks(90)
    got, n = host_sort(dbt, orc, empty, "1")
    assert n == 0 and len(got) == 0
    # too many runs for the chunk: refused, not mangled
    chunked(9)
    with pytest.raises(dbt.DbtError) as e:
        host_sort(dbt, orc, f1, "1")
    assert "chunk is too small" in str(e.value)


def test_out_of_core_long_strings_appear_in_a_late_run(dbt, orc, chunked):
    """Only the last run has strings without a NUL in their first 32 bytes: the resident key columns are
    collected again at the full 120-byte width."""
    f1 = orc.gen_ref(8, 200, two=False)
    rng = np.random.default_rng(11)
    rows = f1["entries"].reshape(-1).copy()
    nlong = 2000                                          # the last 20 blocks
    s = np.zeros((nlong, 120), dtype=np.uint8)
    s[:, :100] = rng.integers(ord("a"), ord("c") + 1, size=(nlong, 100), dtype=np.uint8)
    s[::3, :40] = ord("a")                                # ties through byte 40: decided past the 32-byte prefix
    rows["str"][-nlong:] = s.view("V120").reshape(-1)
    f1["entries"][:] = rows.reshape(f1["entries"].shape)
    chunked(50)
    for field in ("2", "3"):
        got, n = host_sort(dbt, orc, f1, field)
        want = orc.sort(f1, field)
        assert H.same_image(got, want), (field, H.first_diff(got, want))
        assert ooc_stats(dbt)["widened"] == 1
        got, n, u = host_dedup(dbt, orc, f1, field)
        assert H.same_image(got, orc.dedup(f1, field)), field
    # hash join: long strings only in a late S chunk => both sides restart at 120 bytes
    f2 = orc.gen_ref(9, 200, two=False)
    f2["entries"][150:170] = f1["entries"][180:200]
    rc, got, n = host_hashjoin(dbt, orc, f1, f2, "2")
    want = orc.hashjoin(f1, f2, "2")
    assert rc == 0 and n == orc.count_rows(want) and H.same_image(got, want)
    assert ooc_stats(dbt)["widened"] == 1


@pytest.mark.parametrize("field", FIELDS)
def test_out_of_core_hashjoin_streams_s_in_chunks(dbt, orc, chunked, field):
    f1, f2 = orc.gen_ref(9, 150)
    f2["nreserved"][::5] = 77                              # ragged S: matches per chunk are not multiples of 100
    f2["entries"]["valid"][::5, 77:] = 0
    chunked(40)
    rc, got, n = host_hashjoin(dbt, orc, f1, f2, field)
    want = orc.hashjoin(f1, f2, field)
    assert rc == 0 and n == orc.count_rows(want) and H.same_image(got, want), H.first_diff(got, want)
    st = ooc_stats(dbt)
    assert st["runs"] == 4 and st["s_chunks"] == 4


def test_out_of_core_hashjoin_field3_multiplicity_and_capacity(dbt, orc, chunked):
    f1, f2 = orc.gen_ref(12, 30)
    r = f1["entries"].reshape(-1).copy()
    s = f2["entries"].reshape(-1).copy()
    r["num"] = r["num"] % 7                                # few composite keys, many duplicates in R
    r["str"] = r["str"][0]
    s["num"] = s["num"] % 7
    s["str"] = r["str"][0]
    f1["entries"][:] = r.reshape(f1["entries"].shape)
    f2["entries"][:] = s.reshape(f2["entries"].shape)
    want = orc.hashjoin(f1, f2, "3")
    nres = orc.count_rows(want)
    assert nres > 100 * len(f2)                            # every S row matches ~430 R rows
    chunked(10)
    rc, got, n = host_hashjoin(dbt, orc, f1, f2, "3")     # capacity = |S| blocks: too small, the size comes back
    assert rc != 0 and n == nres
    rc, got, n = host_hashjoin(dbt, orc, f1, f2, "3", cap_blocks=H.nb(nres))
    assert rc == 0 and n == nres and H.same_image(got, want)


@pytest.mark.parametrize("field", FIELDS)
def test_out_of_core_mergejoin_is_two_dedups_and_a_streamed_semijoin(dbt, orc, chunked, field):
    L = dbt.lib()
    f1, f2 = orc.gen_ref(9, 150)
    want, wur, wus, info = orc.mergejoin(f1, f2, field)
    for r, s, chunk in ((f1, f2, 40), (f1[:30], f2, 40), (f1, f2[:25], 40)):  # both sides chunked / only one of them
        want, wur, wus, info = orc.mergejoin(r, s, field)
        chunked(chunk)
        res = (C.c_uint64 * 4)()
        o1, o2, o3 = orc.new_blocks(len(r)), orc.new_blocks(len(s)), orc.new_blocks(min(len(r), len(s)))
        dbt.check(L.dbt_host_mergejoin(r.ctypes.data, len(r), s.ctypes.data, len(s), ord(field), o1.ctypes.data, o2.ctypes.data,
                                       o3.ctypes.data, 0, res))
        assert [int(x) for x in res] == [info["nres"], info["nunique_R"], info["nunique_S"], info["later_reads"]], (field, len(r), len(s))
        assert H.same_image(o3[: len(want)], want) and H.same_image(o1[: len(wur)], wur) and H.same_image(o2[: len(wus)], wus)
        # the side images are optional
        dbt.check(L.dbt_host_mergejoin(r.ctypes.data, len(r), s.ctypes.data, len(s), ord(field), None, None, o3.ctypes.data, 0, res))
        assert res[0] == info["nres"] and H.same_image(o3[: len(want)], want)


def test_file_entry_points_switch_to_out_of_core(dbt, orc, chunked, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    f1, f2 = orc.gen_ref(42, 300)
    f1.tofile("file.bin")
    f2.tofile("file2.bin")
    chunked(70)
    name, segs, passes, nios = call_mergesort(dbt, "file.bin", "1", 64)
    want = orc.sort_counters(300, 64)
    assert (segs, passes, nios) == (want["nsorted_segs"], want["npasses"], want["nios"])
    assert H.same_image(read_blocks(orc, name), orc.sort(f1, "1"))
    u, nios = call_dedup(dbt, "file.bin", "3", 64, "nodup.bin")
    want = orc.dedup(f1, "3")
    assert u == orc.count_rows(want) and H.same_image(read_blocks(orc, "nodup.bin"), want)
    n, nios = call_join(dbt, "HashJoin", "file.bin", "file2.bin", "1", 64, "outhash.bin")
    want = orc.hashjoin(f1, f2, "1")
    assert n == orc.count_rows(want) and nios == orc.hashjoin_nios(300, 300, 64, n)
    assert H.same_image(read_blocks(orc, "outhash.bin"), want)
    n, nios = call_join(dbt, "MergeJoin", "file.bin", "file2.bin", "2", 100, "outmerge.bin")
    want, ur, us, info = orc.mergejoin(f1, f2, "2")
    assert n == info["nres"] and nios == orc.mergejoin_nios(300, 300, 100, info)
    assert H.same_image(read_blocks(orc, "outmerge.bin"), want)
    assert H.same_image(read_blocks(orc, "1outfile.bin"), ur) and H.same_image(read_blocks(orc, "2outfile.bin"), us)


def test_out_of_core_jobs_and_in_core_jobs_share_the_slots(dbt, orc, chunked):
    """An out-of-core job finishes inside begin(); wait() then only hands back the counters."""
    L = dbt.lib()
    f1 = orc.gen_ref(4, 120, two=False, num_mod=900)
    out = orc.new_blocks(len(f1))
    chunked(32)
    dbt.check(L.dbt_host_dedup_begin(1, f1.ctypes.data, len(f1), ord("1"), out.ctypes.data, 0))
    chunked(0)
    out2 = orc.new_blocks(len(f1))
    dbt.check(L.dbt_host_dedup_begin(2, f1.ctypes.data, len(f1), ord("1"), out2.ctypes.data, 0))  # in-core
    r = (C.c_uint64 * 4)()
    want = orc.dedup(f1, "1")
    for slot, o in ((1, out), (2, out2)):
        dbt.check(L.dbt_host_job_wait(slot, r))
        assert r[1] == orc.count_rows(want) and H.same_image(o[: len(want)], want)


_OOC_CASES = []
_rr = np.random.default_rng(777)
for _i in range(10):
    _OOC_CASES.append({"seed": 5000 + _i, "recid": str(_rr.choice(["ascending", "shuffled", "duplicates"])),
                       "strlen": str(_rr.choice(["short", "short", "boundary", "long"])), "ragged": bool(_rr.random() < 0.5),
                       "nb": int(_rr.integers(60, 200)), "chunk": int(_rr.integers(32, 64))})


@pytest.mark.parametrize("case", _OOC_CASES, ids=lambda c: f"s{c['seed']}-{c['recid'][:3]}-{c['strlen'][:2]}-{'rag' if c['ragged'] else 'dense'}-{c['nb']}b-c{c['chunk']}")
def test_out_of_core_random_parity(dbt, orc, chunked, case):
    """Random shapes (ragged headers, duplicate recids, strings around the 32-byte boundary, junk after the NUL)
    through every out-of-core operator with a random forced chunk size."""
    from test_gpu_random import random_file

    L = dbt.lib()
    rng = np.random.default_rng(case["seed"])
    r = random_file(orc, rng, case["nb"], case)
    s = random_file(orc, rng, int(rng.integers(40, 160)), case)
    chunked(case["chunk"])
    for field in FIELDS:
        got, n = host_sort(dbt, orc, r, field)
        want = orc.sort(r, field)
        assert H.same_image(got, want), ("sort", field, H.first_diff(got, want))
        got, n, u = host_dedup(dbt, orc, r, field)
        want = orc.dedup(r, field)
        assert H.same_image(got, want), ("dedup", field, H.first_diff(got, want))
        want = orc.hashjoin(r, s, field)
        rc, got, n = host_hashjoin(dbt, orc, r, s, field, cap_blocks=max(H.nb(orc.count_rows(want)), 1))
        assert rc == 0 and H.same_image(got, want), ("hashjoin", field)
        want, wur, wus, info = orc.mergejoin(r, s, field)
        res = (C.c_uint64 * 4)()
        o1, o2, o3 = orc.new_blocks(len(r)), orc.new_blocks(len(s)), orc.new_blocks(min(len(r), len(s)))
        dbt.check(L.dbt_host_mergejoin(r.ctypes.data, len(r), s.ctypes.data, len(s), ord(field), o1.ctypes.data, o2.ctypes.data,
                                       o3.ctypes.data, 0, res))
        assert [int(x) for x in res] == [info["nres"], info["nunique_R"], info["nunique_S"], info["later_reads"]], ("mergejoin", field)
        assert H.same_image(o3[: len(want)], want) and H.same_image(o1[: len(wur)], wur) and H.same_image(o2[: len(wus)], wus)
