"""-m gpu tests of the drop-in boundary: the four dbtproj.h entry points (C++ linkage, called through
their mangled names exactly as main.o would), the host-buffer C-ABI, the committed reference fixtures,
and size-independent properties at a large size."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _entry(dbt, name):
    f = getattr(dbt.lib(), dbt.CXX_ENTRY_POINTS[name])
    f.restype = None
    return f


def call_mergesort(dbt, infile, field, nmem):
    out = C.create_string_buffer(64)
    a, b, c = C.c_uint(), C.c_uint(), C.c_uint()
    _entry(dbt, "MergeSort")(infile.encode(), C.c_ubyte(ord(field)), None, C.c_uint(nmem), out, C.byref(a), C.byref(b), C.byref(c))
    return out.value.decode(), a.value, b.value, c.value


def call_dedup(dbt, infile, field, nmem, outfile):
    a, c = C.c_uint(), C.c_uint()
    _entry(dbt, "EliminateDuplicates")(infile.encode(), C.c_ubyte(ord(field)), None, C.c_uint(nmem), outfile.encode(), C.byref(a), C.byref(c))
    return a.value, c.value


def call_join(dbt, which, in1, in2, field, nmem, outfile):
    a, c = C.c_uint(), C.c_uint()
    _entry(dbt, which)(in1.encode(), in2.encode(), C.c_ubyte(ord(field)), None, C.c_uint(nmem), outfile.encode(), C.byref(a), C.byref(c))
    return a.value, c.value


def read_blocks(orc, path):
    return orc.as_blocks(np.fromfile(path, dtype=np.uint8))


@pytest.fixture()
def workdir(tmp_path, monkeypatch, orc):
    monkeypatch.chdir(tmp_path)
    f1, f2 = orc.gen_ref(42, 300)
    f1.tofile("file.bin")
    f2.tofile("file2.bin")
    return f1, f2


@pytest.mark.parametrize("field,nmem", [("1", 64), ("0", 16), ("2", 8), ("3", 3)])
def test_mergesort_entry_point(dbt, orc, workdir, field, nmem):
    f1, _ = workdir
    name, segs, passes, nios = call_mergesort(dbt, "file.bin", field, nmem)
    want = orc.sort_counters(300, nmem)
    assert (segs, passes, nios) == (want["nsorted_segs"], want["npasses"], want["nios"])
    assert name == f"segment{segs}.bin" and os.path.exists(name)  # outfile is an OUT parameter (DatabaseProject.cpp:375)
    got = read_blocks(orc, name)
    assert H.same_image(got, orc.sort(f1, field)), H.first_diff(got, orc.sort(f1, field))


def test_dedup_hashjoin_mergejoin_entry_points(dbt, orc, workdir):
    f1, f2 = workdir
    u, nios = call_dedup(dbt, "file.bin", "1", 64, "nodup.bin")
    want = orc.dedup(f1, "1")
    assert u == orc.count_rows(want) and nios == orc.dedup_nios(300, 64, u)
    assert H.same_image(read_blocks(orc, "nodup.bin"), want)

    n, nios = call_join(dbt, "HashJoin", "file.bin", "file2.bin", "1", 64, "outhash.bin")
    want = orc.hashjoin(f1, f2, "1")
    assert n == orc.count_rows(want) and nios == orc.hashjoin_nios(300, 300, 64, n)
    assert H.same_image(read_blocks(orc, "outhash.bin"), want)

    n, nios = call_join(dbt, "MergeJoin", "file.bin", "file2.bin", "1", 100, "outmerge.bin")
    want, ur, us, info = orc.mergejoin(f1, f2, "1")
    assert n == info["nres"] and nios == orc.mergejoin_nios(300, 300, 100, info)
    assert H.same_image(read_blocks(orc, "outmerge.bin"), want)
    # the side files main.cpp:121 feeds to HashJoin (DatabaseProject.cpp:385-386)
    assert H.same_image(read_blocks(orc, "1outfile.bin"), ur)
    assert H.same_image(read_blocks(orc, "2outfile.bin"), us)
    # ... and the reference workflow's second step: HashJoin over the deduplicated side files
    n2, _ = call_join(dbt, "HashJoin", "1outfile.bin", "2outfile.bin", "1", 100, "outhash2.bin")
    assert n2 == info["nres"]  # both sides are key-unique: semi-join size == intersection size


def test_entry_points_against_the_live_reference(dbt, orc, workdir):
    if not orc.ref_available():
        pytest.skip("oracle/_ref/ref_runner not built")
    f1, f2 = workdir
    # MergeSort: single merge phase (npasses == 2) => REF is lossless: bit-exact modulo dummy1 after tie canonicalisation
    info, ref_out, _ = orc.run_ref("sort", "1", 301, f1)
    name, segs, passes, nios = call_mergesort(dbt, "file.bin", "1", 301)
    assert (segs, passes) == (info["a"], info["b"]) and name == info["outfile"]
    assert 0 <= info["nios"] - nios <= 301
    refc = orc.rows_of(orc.canonicalise_ties(ref_out[ref_out["nreserved"] > 0], "1")).copy()
    ours = orc.rows_of(read_blocks(orc, name)).copy()
    refc["dummy1"] = 0
    ours["dummy1"] = 0
    assert refc.tobytes() == ours.tobytes()
    # three-pass configuration: counters still equal, REF output is a subsequence of ours
    info, ref_out, _ = orc.run_ref("sort", "1", 8, f1)
    name, segs, passes, nios = call_mergesort(dbt, "file.bin", "1", 8)
    assert (segs, passes) == (info["a"], info["b"]) and passes >= 3
    ours_ids = orc.rows_of(read_blocks(orc, name))["recid"]
    ref_ids = set(orc.rows_of(ref_out)["recid"].tolist())
    assert ref_ids <= set(ours_ids.tolist()) and len(ours_ids) - len(ref_ids) <= 8
    # HashJoin: exact, every field
    for field in "0123":
        info, ref_out, _ = orc.run_ref("hjoin", field, 64, f1, f2)
        n, nios = call_join(dbt, "HashJoin", "file.bin", "file2.bin", field, 64, "oh.bin")
        assert n == info["a"] and abs(info["nios"] - nios) <= 1
        ours = orc.rows_of(read_blocks(orc, "oh.bin"))
        assert np.array_equal(ours["recid"], orc.rows_of(ref_out, info["a"])["recid"])
    # EliminateDuplicates: the reference mis-counts by its documented window
    info, ref_out, _ = orc.run_ref("dedup", "1", 64, f1)
    u, _ = call_dedup(dbt, "file.bin", "1", 64, "nd.bin")
    assert -1 <= info["a"] - u <= 100


def test_reference_error_behaviour(dbt, tmp_path, orc):
    f1 = orc.gen_ref(1, 3, two=False)
    f1.tofile(tmp_path / "file.bin")
    prog = (
        "import ctypes as C, importlib, sys; sys.path.insert(0, %r);"
        "m = importlib.import_module('database-technology-algorithms_b200'); L = m.lib();"
        "f = getattr(L, m.CXX_ENTRY_POINTS['MergeSort']); f.restype = None;"
        "o = C.create_string_buffer(64); a = C.c_uint();"
        "f(b'file.bin', C.c_ubyte(%d), None, C.c_uint(%d), o, C.byref(a), C.byref(a), C.byref(a)); print('RETURNED')"
    )
    p = subprocess.run([sys.executable, "-c", prog % (ROOT, ord("7"), 64)], cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode == 0 and "Wrong field! Please give a field between 0 and 3!" in p.stdout and "RETURNED" not in p.stdout
    p = subprocess.run([sys.executable, "-c", prog % (ROOT, ord("1"), 2)], cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode == 0 and "The buffer size is too small!" in p.stdout and "RETURNED" not in p.stdout
    p = subprocess.run([sys.executable, "-c", (prog % (ROOT, ord("1"), 64)).replace("file.bin", "nofile.bin")], cwd=tmp_path,
                       capture_output=True, text=True)
    assert p.returncode == 1 and "cannot open input file" in p.stderr  # loud, where the reference would segfault


def test_product_reproduces_the_reference_fixtures(dbt, orc):
    meta = json.load(open(os.path.join(HERE, "golden", "ref_golden.json")))
    arr = np.load(os.path.join(HERE, "golden", "ref_golden.npz"))
    f1, f2 = orc.gen_ref(meta["seed"], meta["nblocks"])
    for field in "0123":
        got, n = H.dev_sort(dbt, orc, f1, field)
        assert np.array_equal(orc.rows_of(got)["recid"], arr[f"sort_f{field}"]), field
        got, n = H.dev_hashjoin(dbt, orc, f1, f2, field)
        assert n == meta["counters"][f"hjoin_f{field}"]["nres"]
        assert np.array_equal(orc.rows_of(got)["recid"], arr[f"hjoin_f{field}"]), field
    # field '3' with real multiplicities (REF fixture on a seed where REF does not overflow its output block)
    m = meta["mult"]
    r, s = orc.gen_ref(m["seed"], m["nblocks"], num_mod=m["num_mod"])
    got, n = H.dev_hashjoin(dbt, orc, r, s, "3")
    assert n == m["nres"] and np.array_equal(orc.rows_of(got)["recid"], arr["hjoinm_f3"])


def test_host_scope_operators_with_pageable_and_pinned_buffers(dbt, orc):
    L = dbt.lib()
    f1, f2 = orc.gen_ref(9, 150)
    want = orc.dedup(f1, "3")
    out = orc.new_blocks(len(f1))
    n, u = C.c_uint64(), C.c_uint64()
    dbt.check(L.dbt_host_dedup(f1.ctypes.data, len(f1), ord("3"), out.ctypes.data, 0, C.byref(n), C.byref(u)))  # pageable
    assert u.value == orc.count_rows(want) and H.same_image(out[: len(want)], want)
    # pinned in/out
    nbytes = len(f1) * H.BLOCK_BYTES
    hin, hout = C.c_void_p(), C.c_void_p()
    dbt.check(L.dbt_host_alloc(C.byref(hin), nbytes))
    dbt.check(L.dbt_host_alloc(C.byref(hout), nbytes))
    C.memmove(hin, f2.ctypes.data, nbytes)
    nres = C.c_uint64()
    dbt.check(L.dbt_host_hashjoin(f1.ctypes.data, len(f1), hin, len(f2), ord("1"), hout, len(f2), 0, C.byref(nres)))
    want = orc.hashjoin(f1, f2, "1")
    got = orc.as_blocks(np.ctypeslib.as_array((C.c_uint8 * (len(want) * H.BLOCK_BYTES)).from_address(hout.value)).copy())
    assert nres.value == orc.count_rows(want) and H.same_image(got, want)
    res = (C.c_uint64 * 4)()
    o1, o2, o3 = orc.new_blocks(len(f1)), orc.new_blocks(len(f2)), orc.new_blocks(len(f1))
    dbt.check(L.dbt_host_mergejoin(f1.ctypes.data, len(f1), f2.ctypes.data, len(f2), ord("2"), o1.ctypes.data, o2.ctypes.data,
                                   o3.ctypes.data, 0, res))
    want, wur, wus, info = orc.mergejoin(f1, f2, "2")
    assert [int(x) for x in res] == [info["nres"], info["nunique_R"], info["nunique_S"], info["later_reads"]]
    assert H.same_image(o3[: len(want)], want) and H.same_image(o1[: len(wur)], wur) and H.same_image(o2[: len(wus)], wus)
    out = orc.new_blocks(len(f1))
    dbt.check(L.dbt_host_mergesort(f1.ctypes.data, len(f1), ord("2"), out.ctypes.data, 0, C.byref(n)))
    assert H.same_image(out, orc.sort(f1, "2"))
    L.dbt_host_free(hin)
    L.dbt_host_free(hout)


def test_pipelined_host_jobs_on_several_slots(dbt, orc):
    """begin/wait: jobs on different slots are in flight together (own stream, buffers, workspace) and
    every one lands the oracle's image in its own pinned output."""
    L = dbt.lib()
    assert L.dbt_host_job_slots() >= 2
    inputs = [orc.gen_ref(21 + i, 120 + 40 * i, two=False, num_mod=3000) for i in range(3)]
    pins = []
    for f in inputs:
        nbytes = len(f) * H.BLOCK_BYTES
        hin, hout = C.c_void_p(), C.c_void_p()
        dbt.check(L.dbt_host_alloc(C.byref(hin), nbytes))
        dbt.check(L.dbt_host_alloc(C.byref(hout), nbytes))
        C.memmove(hin, f.ctypes.data, nbytes)
        pins.append((hin, hout, nbytes))

    def image(hout, nblocks):
        return orc.as_blocks(np.ctypeslib.as_array((C.c_uint8 * (nblocks * H.BLOCK_BYTES)).from_address(hout.value)).copy())

    for rnd in range(2):  # second round reuses the slots' cached buffers
        dbt.check(L.dbt_host_dedup_begin(0, pins[0][0], len(inputs[0]), ord("1"), pins[0][1], 0))
        dbt.check(L.dbt_host_mergesort_begin(1, pins[1][0], len(inputs[1]), ord("2"), pins[1][1], 0))
        dbt.check(L.dbt_host_dedup_begin(2, pins[2][0], len(inputs[2]), ord("3"), pins[2][1], 0))
        with pytest.raises(dbt.DbtError) as e:   # a busy slot refuses a second job
            dbt.check(L.dbt_host_dedup_begin(1, pins[0][0], len(inputs[0]), ord("1"), pins[0][1], 0))
        assert "busy" in str(e.value)
        r = (C.c_uint64 * 4)()
        dbt.check(L.dbt_host_job_wait(2, r))
        want = orc.dedup(inputs[2], "3")
        assert (r[0], r[1]) == (orc.count_rows(inputs[2]), orc.count_rows(want))
        assert H.same_image(image(pins[2][1], len(want)), want)
        dbt.check(L.dbt_host_job_wait(0, r))
        want = orc.dedup(inputs[0], "1")
        assert r[1] == orc.count_rows(want) and H.same_image(image(pins[0][1], len(want)), want)
        dbt.check(L.dbt_host_job_wait(1, r))
        assert r[0] == orc.count_rows(inputs[1]) and H.same_image(image(pins[1][1], len(inputs[1])), orc.sort(inputs[1], "2"))
    with pytest.raises(dbt.DbtError):            # nothing in flight any more
        dbt.check(L.dbt_host_job_wait(1, None))
    with pytest.raises(dbt.DbtError):
        dbt.check(L.dbt_host_dedup_begin(99, pins[0][0], 1, ord("1"), pins[0][1], 0))
    # join jobs, then trim and run again from cold buffers
    f1, f2 = orc.gen_ref(9, 100)
    out = orc.new_blocks(len(f2))
    dbt.check(L.dbt_host_hashjoin_begin(3, f1.ctypes.data, len(f1), f2.ctypes.data, len(f2), ord("1"), out.ctypes.data, len(f2), 0))
    o3 = orc.new_blocks(len(f1))
    dbt.check(L.dbt_host_mergejoin_begin(1, f1.ctypes.data, len(f1), f2.ctypes.data, len(f2), ord("1"), None, None, o3.ctypes.data, 0))
    with pytest.raises(dbt.DbtError):
        dbt.check(L.dbt_host_trim())             # refused while jobs are in flight
    r = (C.c_uint64 * 4)()
    dbt.check(L.dbt_host_job_wait(3, r))
    want = orc.hashjoin(f1, f2, "1")
    assert r[0] == orc.count_rows(want) and H.same_image(out[: len(want)], want)
    dbt.check(L.dbt_host_job_wait(1, r))
    want, _, _, info = orc.mergejoin(f1, f2, "1")
    assert r[0] == info["nres"] and H.same_image(o3[: len(want)], want)
    dbt.check(L.dbt_host_trim())
    n, u = C.c_uint64(), C.c_uint64()
    dbt.check(L.dbt_host_dedup(pins[0][0], len(inputs[0]), ord("1"), pins[0][1], 0, C.byref(n), C.byref(u)))
    want = orc.dedup(inputs[0], "1")
    assert u.value == orc.count_rows(want) and H.same_image(image(pins[0][1], len(want)), want)
    for hin, hout, _ in pins:
        L.dbt_host_free(hin)
        L.dbt_host_free(hout)


def test_large_dedup_properties_and_cpu_spot_checks(dbt, orc):
    """20M rows of the bench distribution (generated on the device): size-independent properties checked
    with torch as an independent checker, plus oracle spot checks of the generator on sub-ranges."""
    import torch

    n, U = 20_000_000, 18_000_000
    nb = n // 100
    d_in = H.dev_alloc(nb * H.BLOCK_BYTES)
    d_out = H.dev_alloc(nb * H.BLOCK_BYTES)
    dbt.check(dbt.lib().dbt_gen_syn(42, n, U, 0, 0, n, 0, d_in.data_ptr(), H.stream()))
    torch.cuda.synchronize()
    # generator == CPU restatement on two sub-ranges
    for row0 in (0, 12_345_600):
        want = orc.gen_syn(42, n, U, 0, row0=row0, nrows=500)
        got = d_in[row0 // 100 * H.BLOCK_BYTES:(row0 // 100 + 5) * H.BLOCK_BYTES].cpu().numpy()
        assert got.tobytes() == want.tobytes()
    wsb = dbt.dev_ws_bytes(dbt.OP_DEDUP, nb, 0, "1")
    ws = H.dev_alloc(wsb)
    rows, uniq = dbt.dev_dedup(d_in.data_ptr(), nb, "1", d_out.data_ptr(), ws.data_ptr(), wsb, H.stream())
    assert (rows, uniq) == (n, U)
    img_in = d_in[: nb * H.BLOCK_BYTES].view(torch.int32).view(nb, 3504)
    recs_in = img_in[:, 2:3502].reshape(nb, 100, 35)
    keys = (recs_in[:, :, 1].reshape(-1).to(torch.int64)) & 0xFFFFFFFF
    ids = recs_in[:, :, 0].reshape(-1).to(torch.int64)
    nbo = U // 100
    img_out = d_out[: nbo * H.BLOCK_BYTES].view(torch.int32).view(nbo, 3504)
    recs_out = img_out[:, 2:3502].reshape(nbo, 100, 35)
    okeys = (recs_out[:, :, 1].reshape(-1).to(torch.int64)) & 0xFFFFFFFF
    oids = recs_out[:, :, 0].reshape(-1).to(torch.int64)
    assert bool((okeys[1:] > okeys[:-1]).all())                       # strictly ascending => sorted and unique
    order = torch.argsort(keys * (1 << 31) + ids)                     # (key, recid) order, recid < 2^31 here
    sk, si = keys[order], ids[order]
    first = torch.ones_like(sk, dtype=torch.bool)
    first[1:] = sk[1:] != sk[:-1]
    assert torch.equal(sk[first], okeys) and torch.equal(si[first], oids)   # the min-recid row of every key
    # rows moved verbatim: compare full 140-byte records for a sample of output rows
    sel = torch.randint(0, U, (2000,), device="cuda")
    src_rows = oids[sel]                                              # recid == input row index for this generator
    a = recs_out.reshape(-1, 35)[sel]
    b = recs_in.reshape(-1, 35)[src_rows]
    assert torch.equal(a, b)
    assert bool((img_out[:, 1] == 100).all()) and bool((img_out[:, 0] == torch.arange(nbo, device="cuda", dtype=torch.int32)).all())


def test_baseline_config1_at_full_size(dbt, orc):
    """BASELINE.json configs[1] as stated -- EliminateDuplicates field=num on 100M records, one GPU -- checked through
    size-independent properties by the independent torch checker that bench.py uses for the multi-GPU runs
    (bench_verify.check_dedup_u32: keys strictly ascending, every emitted row is the min-recid row of its key, every key
    once, and a 64-bit multiset hash over all 140 bytes of the emitted records equal to that of the winning input rows)."""
    torch, V = _fresh_device_memory(80)

    n, U = 100_000_000, 90_000_000
    nb = n // 100
    d_in = H.dev_alloc(nb * H.BLOCK_BYTES)
    d_out = H.dev_alloc(nb * H.BLOCK_BYTES)
    dbt.check(dbt.lib().dbt_gen_syn(42, n, U, 0, 0, n, 0, d_in.data_ptr(), H.stream()))
    wsb = dbt.dev_ws_bytes(dbt.OP_DEDUP, nb, 0, "1")
    ws = H.dev_alloc(wsb)
    rows, uniq = dbt.dev_dedup(d_in.data_ptr(), nb, "1", d_out.data_ptr(), ws.data_ptr(), wsb, H.stream())
    assert (rows, uniq) == (n, U)
    del ws
    torch.cuda.empty_cache()
    res = V.check_dedup_u32(d_in, n, d_out, uniq, 1)
    assert res["keys_strictly_ascending"] and res["every_row_is_min_recid_of_its_key"] and res["record_multiset_hash_equal"], res
    assert res["rows"] == U and res["distinct_keys_in_input"] == U, res
    nbo = U // 100  # CANON headers of the packed output image
    hdr = d_out[: nbo * H.BLOCK_BYTES].view(torch.int32).view(nbo, 3504)
    assert bool((hdr[:, 1] == 100).all()) and bool((hdr[:, 0] == torch.arange(nbo, device="cuda", dtype=torch.int32)).all())
    assert bool((hdr[:, 3502] == 1).all()) and bool((hdr[:, 3503] == 100).all())


@pytest.mark.parametrize("kind,label", [(1, "uniform"), (4, "zipf1.1")])
def test_baseline_config3_shard_at_full_size(dbt, orc, kind, label):
    """BASELINE.json configs[3] (HashJoin field=num, R = 100M x S = 1B, uniform and Zipf(1.1) keys) at the size one of four
    GPUs sees -- all of R, a quarter of S -- checked row by row (all 140 bytes, S file order) by bench_verify's torch
    boolean-table filter; the full S streams through one GPU in bench.py, whose match count must agree with N-GPU runs."""
    torch, V = _fresh_device_memory(120)

    L = dbt.lib()
    nr, ns, D = 100_000_000, 250_000_000, 100_000_000
    nbr, nbs = nr // 100, ns // 100
    cap = int(nbs * 0.78) + 64
    d_r, d_s, d_o = H.dev_alloc(nbr * H.BLOCK_BYTES), H.dev_alloc(nbs * H.BLOCK_BYTES), H.dev_alloc(cap * H.BLOCK_BYTES)
    dbt.check(L.dbt_gen_syn(7, nr, D, 1, 0, nr, 0, d_r.data_ptr(), H.stream()))
    dbt.check(L.dbt_gen_syn(9, 1_000_000_000, D, kind, 0, ns, 0, d_s.data_ptr(), H.stream()))  # the first quarter of the 1B-row S
    wsb = dbt.dev_ws_bytes(dbt.OP_HASHJOIN, nbr, nbs, "1")
    ws = H.dev_alloc(wsb)
    k = dbt.dev_hashjoin(d_r.data_ptr(), nbr, d_s.data_ptr(), nbs, "1", d_o.data_ptr(), cap, ws.data_ptr(), wsb, H.stream())
    del ws
    torch.cuda.empty_cache()
    res = V.check_semijoin_u32(d_r, nr, d_s, ns, d_o, k, 1, D)
    assert res["same_rows_in_s_order"] and res["rows"] == res["rows_expected"] == k, res
    assert 0.6 * ns < k < 0.75 * ns  # selectivity 1 - 1/e for uniform keys over |R| values, a little more under Zipf(1.1)


def _fresh_device_memory(gib):
    import gc

    import torch

    gc.collect()
    torch.cuda.empty_cache()  # (an earlier test's cached blocks are not "free" for mem_get_info)
    if torch.cuda.mem_get_info()[0] < gib * 2**30:
        pytest.skip(f"needs ~{gib} GB of free device memory")
    sys.path.insert(0, ROOT)
    import bench_verify as V

    return torch, V


@pytest.mark.parametrize("field,seed,U", [("2", 77, 1_000_000_000), ("1", 78, 1 << 32)])
def test_baseline_config2_and_north_star_sort_shard_at_full_size(dbt, orc, field, seed, U):
    """BASELINE.json configs[2] (MergeSort field=str, 1B records on 8 GPUs) and north_star's num-key MergeSort at the size
    one of the eight GPUs sorts -- 125M records -- through bench_verify.check_sort: (key, recid) order over the whole output
    and a 64-bit multiset hash over all 140 bytes of every record equal to the input's."""
    torch, V = _fresh_device_memory(70)
    n = 125_000_000
    nb = n // 100
    d_in, d_out = H.dev_alloc(nb * H.BLOCK_BYTES), H.dev_alloc(nb * H.BLOCK_BYTES)
    dbt.check(dbt.lib().dbt_gen_syn(seed, 1_000_000_000, U, 1, 0, n, 0, d_in.data_ptr(), H.stream()))  # rank 0's shard of the 1B rows
    wsb = dbt.dev_ws_bytes(dbt.OP_SORT, nb, 0, field)
    ws = H.dev_alloc(wsb)
    rows = dbt.dev_mergesort(d_in.data_ptr(), nb, field, d_out.data_ptr(), ws.data_ptr(), wsb, H.stream())
    assert rows == n
    del ws
    torch.cuda.empty_cache()
    key_fn = V.str_key64 if field == "2" else (lambda im, m: V.column(im, m, 1))
    res = V.check_sort(d_in, n, d_out, rows, key_fn)
    assert res["ordered_in_rank"] and res["record_multiset_hash_equal"] and res["rows"] == n, res


def test_baseline_config4_shard_at_full_size(dbt, orc):
    """BASELINE.json configs[4] (MergeJoin field=num+str, 2 x 500M records on 8 GPUs) at the size one of the eight GPUs
    joins -- 2 x 62.5M records -- through bench_verify.check_mergejoin_composite: R's min-recid row of every (num, str) key
    present in both relations, ascending, all 140 bytes (expectation from torch.sort / unique / isin of the key columns)."""
    torch, V = _fresh_device_memory(70)
    n = 62_500_000
    nb = n // 100
    U = int(0.3 * n)
    d_r, d_s = H.dev_alloc(nb * H.BLOCK_BYTES), H.dev_alloc(nb * H.BLOCK_BYTES)
    d_ur, d_us, d_o = (H.dev_alloc(nb * H.BLOCK_BYTES) for _ in range(3))
    dbt.check(dbt.lib().dbt_gen_syn(21, n, U, 1, 0, n, 0, d_r.data_ptr(), H.stream()))
    dbt.check(dbt.lib().dbt_gen_syn(21 ^ 0x5EED, n, U, 3, 0, n, 0, d_s.data_ptr(), H.stream()))
    wsb = dbt.dev_ws_bytes(dbt.OP_MERGEJOIN, nb, nb, "3")
    ws = H.dev_alloc(wsb)
    info = dbt.dev_mergejoin(d_r.data_ptr(), nb, d_s.data_ptr(), nb, "3", d_ur.data_ptr(), d_us.data_ptr(), d_o.data_ptr(),
                             ws.data_ptr(), wsb, H.stream())
    del ws, d_ur, d_us
    torch.cuda.empty_cache()
    res = V.check_mergejoin_composite(d_r, n, d_s, n, d_o, info["nres"])
    assert res["same_keys_and_recids_in_global_order"] and res["record_multiset_hash_equal"], res
    assert res["rows"] == res["rows_expected"] == info["nres"] > 0, res


def test_driver_runs_the_reference_workflow(dbt, orc, tmp_path):
    """dbt_main = the reference main.cpp workflow (generate, MergeJoin, HashJoin on the side files)."""
    exe = os.path.join(os.path.dirname(dbt.LIB_PATH), "dbt_main")
    if not os.path.exists(exe):
        pytest.skip("dbt_main not built")
    p = subprocess.run([exe, "--nblocks", "200", "--seed", "42", "--nmem", "100", "--ops", "sort,dedup,mjoin,hjoin", "--keep"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    f1, f2 = orc.gen_ref(42, 200)
    assert H.same_image(read_blocks(orc, tmp_path / "file.bin"), f1)      # the driver's generator == main.cpp:41-77
    assert H.same_image(read_blocks(orc, tmp_path / "file2.bin"), f2)
    want, ur, us, info = orc.mergejoin(f1, f2, "1")
    assert H.same_image(read_blocks(orc, tmp_path / "outmerge.bin"), want)
    assert f"pairs in the output {info['nres']} " in p.stdout
    c = orc.sort_counters(200, 100)
    assert os.path.exists(tmp_path / f"segment{c['nsorted_segs']}.bin")
    assert H.same_image(read_blocks(orc, tmp_path / "NOduplicates.bin"), orc.dedup(f1, "1"))


def test_hashjoin_field3_output_many_times_larger_than_s(dbt, orc, tmp_path, monkeypatch):
    """Field '3' emits an S row once per matching R row (DatabaseProject.cpp:616-629): with ~10 equal (num, str) rows per
    key in R the output is ~10x S.  The entry point retries with the exact capacity, and the workspace must grow with it."""
    monkeypatch.chdir(tmp_path)
    nb = 10_000                                     # 1M rows per side
    r = orc.gen_syn(5, nb * 100, 1000, 1)           # num uniform over 1000 values
    s = orc.gen_syn(6, nb * 100, 1000, 1)
    for img in (r, s):                              # one string everywhere: composite key == num, ~1000 rows per key in R ...
        e = img["entries"]
        e["str"] = np.zeros(120, np.uint8).view("V120")[0]
    e = r["entries"].reshape(-1).copy()
    e["num"][100_000:] += 5000                      # ... but only R's first 100,000 rows can match: ~100 R rows per key
    r["entries"][:] = e.reshape(r["entries"].shape)
    e = s["entries"].reshape(-1).copy()
    e["num"][100_000:] += 9000                      # and only S's first 100,000 rows probe them: ~10M output rows = 10 x S
    s["entries"][:] = e.reshape(s["entries"].shape)
    r.tofile("r.bin")
    s.tofile("s.bin")
    n, nios = call_join(dbt, "HashJoin", "r.bin", "s.bin", "3", 64, "out.bin")
    rn = orc.rows_of(r)["num"][:100_000]
    sn = orc.rows_of(s)["num"][:100_000]
    per_key = np.bincount(rn, minlength=1000)
    want_n = int(per_key[sn].sum())
    assert want_n > 9 * nb * 100 and n == want_n
    out = read_blocks(orc, "out.bin")
    assert orc.count_rows(out) == want_n
    got_ids = orc.rows_of(out)["recid"]
    want_ids = np.repeat(orc.rows_of(s)["recid"][:100_000], per_key[sn])
    assert np.array_equal(got_ids, want_ids)        # S file order, each row repeated once per matching R row
    assert nios == orc.hashjoin_nios(nb, nb, 64, n)


def test_hashjoin_with_one_memory_block_mirrors_the_reference(dbt, orc, workdir):
    """nmem_blocks == 1: the reference reads 0 blocks per fread and stops at once in both phases (DatabaseProject.cpp:521-525,
    564-568): nres = 0, two counted reads, an empty output file."""
    n, nios = call_join(dbt, "HashJoin", "file.bin", "file2.bin", "1", 1, "oh1.bin")
    assert (n, nios) == (0, 2) and os.path.getsize("oh1.bin") == 0
    if orc.ref_available():
        f1, f2 = workdir
        info, ref_out, _ = orc.run_ref("hjoin", "1", 1, f1, f2)
        assert info["a"] == 0 and info["nios"] == 2


def test_config0_mergesort_file_of_1m_records_through_the_entry_point(dbt, orc, tmp_path, monkeypatch):
    """BASELINE configs[0] through the real drop-in call: a 1M-record file (140 MB: several chunks of the file <-> pinned
    <-> device pipeline in both directions), field num, nmem_blocks = 64.  Image bit-exact, counters = SURVEY Appendix B."""
    monkeypatch.chdir(tmp_path)
    f1 = orc.gen_ref(42, 10000, two=False)
    f1.tofile("file.bin")
    name, segs, passes, nios = call_mergesort(dbt, "file.bin", "1", 64)
    assert (segs, passes, nios) == (161, 3, 30048) and name == "segment161.bin"
    got = read_blocks(orc, name)
    want = orc.sort(f1, "1")
    assert H.same_image(got, want), H.first_diff(got, want)
    u, nios = call_dedup(dbt, "file.bin", "1", 64, "nodup.bin")
    want = orc.dedup(f1, "1")
    assert u == orc.count_rows(want) and H.same_image(read_blocks(orc, "nodup.bin"), want)
