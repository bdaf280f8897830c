"""-m gpu parity tests: the CUDA path through the C-ABI vs the CANON oracle, bit-exact."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

FIELDS = ["0", "1", "2", "3"]


@pytest.fixture(scope="module")
def files(orc):
    f1, f2 = orc.gen_ref(42, 300)          # 30,000 rows, reference generator (main.cpp:41-77)
    g1, g2 = orc.gen_ref(7, 120, tail_mode=1)  # junk after the NUL / in dummies, as the reference leaves it
    return {"f1": f1, "f2": f2, "g1": g1, "g2": g2}


@pytest.mark.parametrize("field", FIELDS)
@pytest.mark.parametrize("name", ["f1", "g2"])
def test_mergesort_matches_canon(dbt, orc, files, field, name):
    src = files[name]
    got, n = H.dev_sort(dbt, orc, src, field)
    want = orc.sort(src, field)
    assert n == orc.count_rows(src)
    assert H.same_image(got, want), H.first_diff(got, want)


@pytest.mark.parametrize("field", FIELDS)
def test_dedup_matches_canon(dbt, orc, files, field):
    src = files["f1"]
    got, n, u = H.dev_dedup(dbt, orc, src, field)
    want = orc.dedup(src, field)
    assert u == orc.count_rows(want)
    assert H.same_image(got, want), H.first_diff(got, want)


@pytest.mark.parametrize("field", FIELDS)
def test_hashjoin_matches_canon(dbt, orc, files, field):
    r, s = files["f1"], files["f2"]
    got, n = H.dev_hashjoin(dbt, orc, r, s, field)
    want = orc.hashjoin(r, s, field)
    assert n == orc.count_rows(want)
    assert H.same_image(got, want), H.first_diff(got, want)


def test_hashjoin_field3_multiplicity(dbt, orc):
    # small num domain => many (num,str) duplicates in R ("Hola" rows): S rows are emitted once per matching R row
    r, s = orc.gen_ref(3, 40, num_mod=3)
    want = orc.hashjoin(r, s, "3")
    nres = orc.count_rows(want)
    assert nres > orc.count_rows(s) // 100  # really has multiplicities
    got, n = H.dev_hashjoin(dbt, orc, r, s, "3", cap_blocks=H.nb(nres))
    assert n == nres
    assert H.same_image(got, want), H.first_diff(got, want)
    # capacity too small: loud error carrying the needed size
    with pytest.raises(dbt.DbtError):
        H.dev_hashjoin(dbt, orc, r, s, "3", cap_blocks=1)


@pytest.mark.parametrize("field", FIELDS)
def test_mergejoin_matches_canon(dbt, orc, files, field):
    r, s = files["f1"], files["f2"]
    got, ur, us, info = H.dev_mergejoin(dbt, orc, r, s, field)
    want, wur, wus, winfo = orc.mergejoin(r, s, field)
    assert info == winfo
    assert H.same_image(ur, wur), H.first_diff(ur, wur)
    assert H.same_image(us, wus), H.first_diff(us, wus)
    assert H.same_image(got, want), H.first_diff(got, want)


def test_config1_sort_1m_rows(dbt, orc):
    # BASELINE config[0]: 1M records, field num (nmem_blocks only affects the counters)
    f1 = orc.gen_ref(42, 10000, two=False)
    got, n = H.dev_sort(dbt, orc, f1, "1")
    want = orc.sort(f1, "1")
    assert n == 1_000_000
    assert H.same_image(got, want), H.first_diff(got, want)


def test_ragged_and_empty_inputs(dbt, orc):
    f1 = orc.gen_ref(11, 50, two=False)
    rag = f1.copy()
    rng = np.random.default_rng(5)
    rag["nreserved"] = rng.integers(0, 101, size=len(rag)).astype(np.uint32)  # partial blocks anywhere, some empty
    for field in FIELDS:
        got, n = H.dev_sort(dbt, orc, rag, field)
        want = orc.sort(rag, field)
        assert n == orc.count_rows(rag)
        assert H.same_image(got, want), (field, H.first_diff(got, want))
    got, n, u = H.dev_dedup(dbt, orc, rag, "1")
    assert H.same_image(got, orc.dedup(rag, "1"))
    # zero blocks, and blocks with zero rows
    empty = orc.new_blocks(0)
    got, n = H.dev_sort(dbt, orc, empty, "1")
    assert n == 0 and len(got) == 0
    hollow = f1[:3].copy()
    hollow["nreserved"] = 0
    got, n, u = H.dev_dedup(dbt, orc, hollow, "2")
    assert (n, u, len(got)) == (0, 0, 0)
    got, n = H.dev_hashjoin(dbt, orc, hollow, f1, "1")
    assert n == 0
    got, n = H.dev_hashjoin(dbt, orc, f1, hollow, "1")
    assert n == 0


def test_ties_canonicalised_by_recid_when_file_order_is_not(dbt, orc):
    f1 = orc.gen_ref(13, 60, two=False, num_mod=50)  # heavy key duplication
    rows = f1["entries"].reshape(-1)
    rng = np.random.default_rng(1)
    rows[:] = rows[rng.permutation(len(rows))]       # recids no longer ascending in file order
    for field in FIELDS:
        got, n = H.dev_sort(dbt, orc, f1, field)
        want = orc.sort(f1, field)
        assert H.same_image(got, want), (field, H.first_diff(got, want))
        got, n, u = H.dev_dedup(dbt, orc, f1, field)
        assert H.same_image(got, orc.dedup(f1, field)), field


def test_long_strings_use_full_120_byte_keys(dbt, orc):
    f1 = orc.gen_ref(17, 20, two=False)
    rows = f1["entries"].reshape(-1)
    rng = np.random.default_rng(2)
    s = np.zeros((len(rows), 120), dtype=np.uint8)
    lens = rng.integers(0, 120, size=len(rows))
    for i, L in enumerate(lens):   # common 40-byte prefix so that the order is decided past byte 32
        s[i, :L] = np.concatenate([np.full(40, ord("x")), rng.integers(97, 123, size=80)]).astype(np.uint8)[:L]
    rows["str"] = s.view("V120").reshape(-1)
    for field in ["2", "3"]:
        d_in = H.to_dev(f1)
        d_out = H.dev_alloc(len(f1) * H.BLOCK_BYTES)
        wsb = dbt.dev_ws_bytes(dbt.OP_DEDUP, len(f1), 0, field, 30)
        ws = H.dev_alloc(wsb)
        n, u = dbt.dev_dedup(d_in.data_ptr(), len(f1), field, d_out.data_ptr(), ws.data_ptr(), wsb, H.stream())
        got = H.to_host(d_out, H.nb(u), orc)
        want = orc.dedup(f1, field)
        assert H.same_image(got, want), (field, H.first_diff(got, want))


def test_wrong_field_is_an_error(dbt, orc):
    f1 = orc.gen_ref(1, 2, two=False)
    with pytest.raises(dbt.DbtError):
        H.dev_sort(dbt, orc, f1, "7")


def test_sort_pairs_u32_full_range(dbt):
    import ctypes as C
    import torch

    n = 3_000_017
    g = torch.Generator(device="cuda").manual_seed(3)
    keys = torch.randint(0, 2**32, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.uint32)
    vals = torch.arange(n, dtype=torch.int32, device="cuda")
    k1, v1 = keys.clone(), vals.clone()
    k2, v2 = torch.empty_like(k1), torch.empty_like(v1)
    wsb = dbt.lib().dbt_sort_pairs_ws_bytes(n)
    ws = H.dev_alloc(wsb)
    alt = C.c_int()
    dbt.check(dbt.lib().dbt_sort_pairs_u32(k1.data_ptr(), k2.data_ptr(), v1.data_ptr(), v2.data_ptr(), n, 0, 32,
                                           ws.data_ptr(), wsb, H.stream(), C.byref(alt)))
    ko, vo = (k2, v2) if alt.value else (k1, v1)
    ref_k, ref_i = torch.sort(keys.to(torch.int64), stable=True)
    assert torch.equal(ko.to(torch.int64), ref_k)
    assert torch.equal(vo.to(torch.int64), ref_i)


def test_sort_pairs_beyond_2_30_elements_uses_wide_tile_states(dbt):
    """n >= 2^30: the tile states of the look-back chain switch to 64-bit words (32-bit words hold 30-bit positions).
    Checked by properties, in chunks: (key, value) strictly ascending, so the values are distinct, i.e. a permutation,
    ties in input order (stable); and every output key is the input key of its value."""
    import ctypes as C
    import torch

    if torch.cuda.mem_get_info()[1] < 60 * 2**30:
        pytest.skip("needs a 60 GB device")
    n = 2**30 + 77_777
    g = torch.Generator(device="cuda").manual_seed(9)
    keys = torch.randint(-2**31, 2**31, (n,), dtype=torch.int32, device="cuda", generator=g)
    k1 = keys.clone()
    v1 = torch.arange(n, dtype=torch.int32, device="cuda")
    k2, v2 = torch.empty_like(k1), torch.empty_like(v1)
    wsb = dbt.lib().dbt_sort_pairs_ws_bytes(n)
    ws = H.dev_alloc(wsb)
    alt = C.c_int()
    dbt.check(dbt.lib().dbt_sort_pairs_u32(k1.data_ptr(), k2.data_ptr(), v1.data_ptr(), v2.data_ptr(), n, 0, 32,
                                           ws.data_ptr(), wsb, H.stream(), C.byref(alt)))
    ko, vo = (k2, v2) if alt.value else (k1, v1)
    step = 2**27
    for a in range(0, n, step):
        b = min(n, a + step + 1)  # one element of overlap: the pair across the chunk boundary is checked too
        kk = ko[a:b].to(torch.int64) & 0xFFFFFFFF
        vv = vo[a:b].to(torch.int64)
        assert bool(((vv >= 0) & (vv < n)).all())
        asc = (kk[1:] > kk[:-1]) | ((kk[1:] == kk[:-1]) & (vv[1:] > vv[:-1]))
        assert bool(asc.all()), f"order broken in chunk at {a}"
        assert bool((keys[vv] == ko[a:b]).all()), f"pairs broken in chunk at {a}"


def test_hashjoin_wide_key_range_and_table_fallback(dbt, orc, monkeypatch):
    """u32 semi-join paths: L2-resident bitmap (narrow key range, the other tests), sliced full-range bitmap
    (keys spread over 32 bits), and the linear-probing hash table (forced)."""
    r, s = orc.gen_ref(21, 150)
    rng = np.random.default_rng(8)
    pool = rng.integers(0, 2**32, size=9000, dtype=np.uint64).astype(np.uint32)  # shared pool => plenty of matches
    pool[:3] = [0, 0xFFFFFFFF, 0x80000000]                                        # extremes, incl. the table's sentinel value
    for img in (r, s):
        e = img["entries"]
        e["num"] = pool[rng.integers(0, len(pool), size=e["num"].shape)]
    want = orc.hashjoin(r, s, "1")
    assert 0 < orc.count_rows(want) < orc.count_rows(s)
    got, n = H.dev_hashjoin(dbt, orc, r, s, "1")          # two streaming passes over S against the full-range bitmap (default)
    assert H.same_image(got, want), H.first_diff(got, want)
    monkeypatch.setenv("DBT_JOIN_FUSED", "1")
    got, n = H.dev_hashjoin(dbt, orc, r, s, "1")          # one chained streaming pass
    assert H.same_image(got, want), H.first_diff(got, want)
    monkeypatch.setenv("DBT_JOIN_FUSED", "0")              # the column-based paths:
    got, n = H.dev_hashjoin(dbt, orc, r, s, "1")          # sliced bitmap
    assert H.same_image(got, want), H.first_diff(got, want)
    monkeypatch.setenv("DBT_JOIN_NO_SLICES", "1")          # -> hash table (span too wide for the single bitmap)
    got, n = H.dev_hashjoin(dbt, orc, r, s, "1")
    assert H.same_image(got, want), H.first_diff(got, want)
    monkeypatch.setenv("DBT_JOIN_NO_BITMAP", "1")
    r2, s2 = orc.gen_ref(22, 80)
    got, n = H.dev_hashjoin(dbt, orc, r2, s2, "1")        # narrow range, table forced
    assert H.same_image(got, orc.hashjoin(r2, s2, "1"))


def test_joins_with_long_strings_on_one_side_only(dbt, orc):
    r, s = orc.gen_ref(31, 12)
    rows = s["entries"].reshape(-1).copy()
    long = np.zeros(120, np.uint8)
    long[:60] = ord("z")
    rows["str"][5] = long.view("V120")[0]              # S has one 60-byte string, R none
    rows["str"][7] = r["entries"].reshape(-1)["str"][3]  # and a guaranteed match
    s["entries"] = rows.reshape(len(s), 100)
    for field in "23":
        want = orc.hashjoin(r, s, field)
        got, n = H.dev_hashjoin(dbt, orc, r, s, field, kw=30)
        assert H.same_image(got, want), (field, H.first_diff(got, want))
        want = orc.hashjoin(s, r, field)
        got, n = H.dev_hashjoin(dbt, orc, s, r, field, kw=30)
        assert H.same_image(got, want), (field, H.first_diff(got, want))
        got, ur, us, info = H.dev_mergejoin(dbt, orc, r, s, field, kw=30)
        want, wur, wus, winfo = orc.mergejoin(r, s, field)
        assert info == winfo and H.same_image(got, want) and H.same_image(us, wus), field


@pytest.mark.parametrize("field", FIELDS)
def test_innerjoin_pairs_match_the_oracle(dbt, orc, files, field):
    import ctypes as C
    import torch

    r, s = orc.gen_ref(5, 60, num_mod=700)  # ~8.6 rows per key: real multiplicities on both sides
    want = orc.innerjoin_pairs(r, s, field)
    d_r, d_s = H.to_dev(r), H.to_dev(s)
    cap = max(len(want), 1)
    d_pairs = torch.empty(cap * 2, dtype=torch.int32, device="cuda")
    wsb = dbt.dev_ws_bytes(dbt.OP_MERGEJOIN, len(r), len(s), field)
    ws = H.dev_alloc(wsb)
    n = C.c_uint64()
    dbt.check(dbt.lib().dbt_dev_innerjoin_pairs(d_r.data_ptr(), len(r), d_s.data_ptr(), len(s), ord(field), d_pairs.data_ptr(), cap,
                                                ws.data_ptr(), wsb, H.stream(), C.byref(n)))
    assert n.value == len(want)
    got = d_pairs.cpu().numpy().view(np.uint32).reshape(-1, 2)[: n.value]
    assert np.array_equal(got, want)  # same order too: S file order, then R rows by (key, recid)
    if len(want) > 1:
        rc = dbt.lib().dbt_dev_innerjoin_pairs(d_r.data_ptr(), len(r), d_s.data_ptr(), len(s), ord(field), d_pairs.data_ptr(), 1,
                                               ws.data_ptr(), wsb, H.stream(), C.byref(n))
        assert rc == -3 and n.value == len(want)  # capacity too small: loud, with the needed size


def test_device_generator_matches_the_cpu_restatement_for_every_kind(dbt, orc):
    import torch

    n, U = 30000, 9000
    for kind, seed in ((0, 42), (1, 7), (2, 9), (3, 21 ^ 0x5EED), (4, 9)):  # 4 = Zipf(1.1): double arithmetic, same bits
        d = H.dev_alloc(n // 100 * H.BLOCK_BYTES)
        dbt.check(dbt.lib().dbt_gen_syn(seed, n, U, kind, 0, n, 0, d.data_ptr(), H.stream()))
        torch.cuda.synchronize()
        got = H.to_host(d, n // 100, orc)
        want = orc.gen_syn(seed, n, U, kind)
        assert H.same_image(got, want), (kind, H.first_diff(got, want))
    r = orc.gen_syn(21, n, U, 1)
    s = orc.gen_syn(21 ^ 0x5EED, n, U, 3)
    matched = orc.count_rows(orc.hashjoin(r, s, "3"))
    assert 0.35 * n < matched < 0.9 * n  # about half of S's rows carry a composite key of R (plus multiplicities)


def test_sort_pairs_unaligned_buffers_take_the_fallback_kernel(dbt):
    """cp.async.bulk needs 16-byte aligned sources; buffers that are only 4-byte aligned go through the
    one-tile-per-CTA onesweep kernel instead.  Same result either way (also covers a partial last tile)."""
    import ctypes as C
    import torch

    n = 1_234_567
    g = torch.Generator(device="cuda").manual_seed(11)
    base = torch.randint(0, 2**32, (n + 8,), dtype=torch.int64, device="cuda", generator=g).to(torch.uint32)
    keys = base[1:n + 1]                                  # data_ptr is 4 bytes past a 16-byte boundary
    assert keys.data_ptr() % 16 == 4
    k1 = torch.empty(n + 8, dtype=torch.uint32, device="cuda")[1:n + 1]
    k1.copy_(keys)
    k2 = torch.empty(n + 8, dtype=torch.uint32, device="cuda")[1:n + 1]
    v1 = torch.arange(n + 8, dtype=torch.int32, device="cuda")[1:n + 1].clone()
    v1b = torch.empty(n + 8, dtype=torch.int32, device="cuda")[1:n + 1]
    v1b.copy_(torch.arange(n, dtype=torch.int32, device="cuda"))
    v2 = torch.empty(n + 8, dtype=torch.int32, device="cuda")[1:n + 1]
    wsb = dbt.lib().dbt_sort_pairs_ws_bytes(n)
    ws = H.dev_alloc(wsb)
    alt = C.c_int()
    dbt.check(dbt.lib().dbt_sort_pairs_u32(k1.data_ptr(), k2.data_ptr(), v1b.data_ptr(), v2.data_ptr(), n, 0, 32, ws.data_ptr(), wsb,
                                           H.stream(), C.byref(alt)))
    ko, vo = (k2, v2) if alt.value else (k1, v1b)
    ref_k, ref_i = torch.sort(keys.to(torch.int64), stable=True)
    assert torch.equal(ko.to(torch.int64), ref_k) and torch.equal(vo.to(torch.int64), ref_i)


def test_misaligned_images_are_rejected_not_faulted(dbt, orc):
    f1 = orc.gen_ref(1, 4, two=False)
    raw = H.dev_alloc(len(f1) * H.BLOCK_BYTES + 64)
    src = H.to_dev(f1)
    mis = raw[4:4 + len(f1) * H.BLOCK_BYTES]
    mis.copy_(src[: len(f1) * H.BLOCK_BYTES])
    out = H.dev_alloc(len(f1) * H.BLOCK_BYTES)
    wsb = dbt.dev_ws_bytes(dbt.OP_SORT, len(f1), 0, "1")
    ws = H.dev_alloc(wsb)
    with pytest.raises(dbt.DbtError) as e:
        dbt.dev_mergesort(mis.data_ptr(), len(f1), "1", out.data_ptr(), ws.data_ptr(), wsb, H.stream())
    assert e.value.code == -1 and "aligned" in str(e.value)
    got, n = H.dev_sort(dbt, orc, f1, "1")  # and the context is still healthy afterwards
    assert H.same_image(got, orc.sort(f1, "1"))


@pytest.mark.parametrize("shape", ["full_range", "shifted3", "high_byte_only", "two_middle_bytes"])
def test_digit_plans_with_and_without_the_extraction_histogram(dbt, orc, shape):
    """Byte-aligned digit plans reuse the histogram made while the keys were extracted; plans whose
    digits straddle byte boundaries count the keys again.  Both must order rows like the oracle."""
    f1 = orc.gen_ref(5, 300, two=False)
    rows = f1["entries"].reshape(-1)
    rng = np.random.default_rng(7)
    raw = rng.integers(0, 1 << 32, size=len(rows), dtype=np.uint64)
    if shape == "full_range":
        num = raw
    elif shape == "shifted3":                       # varying bits 3..14: digits start at bit 3 and 11
        num = (raw & 0xFFF) << 3 | 0x5
    elif shape == "high_byte_only":                 # one aligned pass, on byte 3
        num = (raw & 0xFF) << 24 | 0x00ABCDEF
    else:                                           # bytes 1 and 2 vary, byte 0 and 3 constant
        num = (raw & 0xFFFF) << 8 | 0x7F000011
    rows["num"] = num.astype(np.uint32)
    f1["entries"][:] = rows.reshape(f1["entries"].shape)
    f1["nreserved"][-1] = 37                        # ragged tail block
    f1["entries"]["valid"][-1, 37:] = 0
    for field in ("0", "1"):
        got, n = H.dev_sort(dbt, orc, f1, field)
        want = orc.sort(f1, field)
        assert n == orc.count_rows(f1) and H.same_image(got, want), (shape, field, H.first_diff(got, want))
        got, n, u = H.dev_dedup(dbt, orc, f1, field)
        assert H.same_image(got, orc.dedup(f1, field)), (shape, field)


@pytest.mark.parametrize("shape", ["one_word", "two_words", "too_wide"])
def test_multiword_keys_sorted_through_their_varying_bits(dbt, orc, shape):
    """str / num+str keys whose varying bits fit 32 or 64 bits are sorted (and deduplicated) as the concatenation of
    those bits; wider keys take the word-by-word LSD.  All three must give the oracle's image."""
    f1 = orc.gen_ref(23, 80, two=False, num_mod=7)
    rows = f1["entries"].reshape(-1).copy()
    n = len(rows)
    rng = np.random.default_rng(5)
    s = np.zeros((n, 120), dtype=np.uint8)
    if shape == "one_word":        # 5 letters from 'a'..'d' + one "Hola"-like outlier: ~3 bits x 5 bytes
        s[:, :5] = rng.integers(ord("a"), ord("e"), size=(n, 5), dtype=np.uint8)
    elif shape == "two_words":     # 9 letters a..z: ~5 bits x 9 bytes = 45 bits (+ 3 of num for field '3')
        s[:, :9] = rng.integers(ord("a"), ord("z") + 1, size=(n, 9), dtype=np.uint8)
    else:                          # 20 letters: > 64 varying bits
        s[:, :20] = rng.integers(ord("a"), ord("z") + 1, size=(n, 20), dtype=np.uint8)
    s[::50, 2:] = 0                # some shorter strings: the NUL takes part in the order
    s[1::97] = s[0]                # exact duplicates of one key
    rows["str"] = s.view("V120").reshape(-1)
    rows = rows[rng.permutation(n)]   # recids not in file order: the recid word is sorted below the compact key
    f1["entries"][:] = rows.reshape(f1["entries"].shape)
    f1["nreserved"][3] = 41
    f1["entries"]["valid"][3, 41:] = 0
    for field in ("2", "3"):
        got, cnt = H.dev_sort(dbt, orc, f1, field)
        want = orc.sort(f1, field)
        assert cnt == orc.count_rows(f1) and H.same_image(got, want), (shape, field, H.first_diff(got, want))
        got, cnt, u = H.dev_dedup(dbt, orc, f1, field)
        want = orc.dedup(f1, field)
        assert u == orc.count_rows(want) and H.same_image(got, want), (shape, field, H.first_diff(got, want))


def test_zipf_generator_is_zipf_and_reproducible_on_any_subrange(dbt, orc):
    """kind 4: rank ~ Zipf(s = 1.1) over [1, U] by rejection-inversion, key = bijection(rank).  The device generator and
    the CPU restatement must agree bit for bit on any sub-range of a large relation, and the head must carry Zipf's mass."""
    import torch

    n_total, U = 50_000_000, 100_000_000
    for row0, nrows in ((0, 20000), (31_415_900, 20000), (n_total - 20000, 20000)):
        d = H.dev_alloc(nrows // 100 * H.BLOCK_BYTES)
        dbt.check(dbt.lib().dbt_gen_syn(9, n_total, U, 4, row0, nrows, 0, d.data_ptr(), H.stream()))
        torch.cuda.synchronize()
        got = H.to_host(d, nrows // 100, orc)
        want = orc.gen_syn(9, n_total, U, 4, row0, nrows)
        assert H.same_image(got, want), (row0, H.first_diff(got, want))
    # the most frequent key of a 1M-row sample is rank 1's, with P = 1 / H_{U,1.1} = 0.111 (U = 1e8)
    nrows = 1_000_000
    d = H.dev_alloc(nrows // 100 * H.BLOCK_BYTES)
    dbt.check(dbt.lib().dbt_gen_syn(9, n_total, U, 4, 0, nrows, 0, d.data_ptr(), H.stream()))
    img = d[: nrows // 100 * H.BLOCK_BYTES].view(torch.int32).view(-1, 3504)[:, 2:3502].reshape(-1, 35)
    num = img[:, 1].to(torch.int64) & 0xFFFFFFFF
    vals, counts = torch.unique(num, return_counts=True)
    top = counts.max().item() / nrows
    assert 0.105 < top < 0.118, top
    assert int(vals[counts.argmax()]) == (0 * 2654435761 + 40503) % U  # rho(rank 1 - 1)


def test_fused_semijoin_edge_cases(dbt, orc, monkeypatch):
    """Fields '0'/'1' run HashJoin's probe as streaming passes over S (DBT_JOIN_FUSED=2, the default: count pass + copy
    pass; 1: one chained pass): ragged S blocks (the passes read nreserved from the block they stream), results that end
    exactly on a block boundary, empty results, capacity errors carrying the needed size, agreement with the column path."""
    r, s = orc.gen_ref(77, 240, num_mod=4000)
    rng = np.random.default_rng(3)
    rag = s.copy()
    rag["nreserved"] = rng.integers(0, 101, size=len(rag)).astype(np.uint32)
    rag["nreserved"][::7] = 0
    rag["nreserved"][5] = 100
    for mode in ("2", "1"):
        monkeypatch.setenv("DBT_JOIN_FUSED", mode)
        for field in "01":
            for src in (s, rag):
                want = orc.hashjoin(r, src, field)
                got, n = H.dev_hashjoin(dbt, orc, r, src, field)
                assert n == orc.count_rows(want) and H.same_image(got, want), (mode, field, H.first_diff(got, want))
    monkeypatch.delenv("DBT_JOIN_FUSED")
    # exactly k * 100 matches: no partial last block to fix up
    want = orc.hashjoin(r, s, "1")
    k = orc.count_rows(want)
    ids = orc.rows_of(want)["recid"]
    cut = ids[(k // 100) * 100 - 1]            # keep S rows up to the row that completes the last full block
    srows = s["entries"].reshape(-1)
    keep = srows["recid"] <= cut
    s2 = orc.new_blocks(len(s))
    s2[:] = s
    e2 = s2["entries"].reshape(-1).copy()
    e2["num"][~keep] = 0xFFFFFFF0             # rows past the cut no longer match anything in R
    s2["entries"][:] = e2.reshape(s2["entries"].shape)
    want2 = orc.hashjoin(r, s2, "1")
    assert orc.count_rows(want2) == (k // 100) * 100
    got, n = H.dev_hashjoin(dbt, orc, r, s2, "1")
    assert n == orc.count_rows(want2) and H.same_image(got, want2), H.first_diff(got, want2)
    # nothing matches
    e2["num"][:] = 0xFFFFFFF0
    s2["entries"][:] = e2.reshape(s2["entries"].shape)
    got, n = H.dev_hashjoin(dbt, orc, r, s2, "1")
    assert n == 0 and len(got) == 0
    # capacity too small: loud, with the needed size; the exact capacity then works
    import ctypes as C
    d_r, d_s = H.to_dev(r), H.to_dev(s)
    wsb = dbt.dev_ws_bytes(dbt.OP_HASHJOIN, len(r), len(s), "1")
    ws = H.dev_alloc(wsb)
    d_out = H.dev_alloc(H.nb(k) * H.BLOCK_BYTES)
    nres = C.c_uint64()
    rc = dbt.lib().dbt_dev_hashjoin(d_r.data_ptr(), len(r), d_s.data_ptr(), len(s), ord("1"), d_out.data_ptr(), 2, ws.data_ptr(), wsb,
                                    H.stream(), C.byref(nres))
    assert rc == -3 and nres.value == k
    rc = dbt.lib().dbt_dev_hashjoin(d_r.data_ptr(), len(r), d_s.data_ptr(), len(s), ord("1"), d_out.data_ptr(), H.nb(k), ws.data_ptr(),
                                    wsb, H.stream(), C.byref(nres))
    assert rc == 0 and H.same_image(H.to_host(d_out, H.nb(k), orc), want)
    # the column-based path gives the same image
    monkeypatch.setenv("DBT_JOIN_FUSED", "0")
    got, n = H.dev_hashjoin(dbt, orc, r, rag, "0")
    assert H.same_image(got, orc.hashjoin(r, rag, "0"))


def test_radix_partitioned_join_agrees_with_the_hbm_table(dbt, orc, monkeypatch):
    """Fields '2' (more than 32 varying key bits) and '3' take the radix-partitioned build / probe with per-partition tables
    in shared memory when the keys' varying bits fit 64; DBT_JOIN_NO_RADIX=1 forces the linear-probing table in HBM.
    Both must give the oracle's image: set semantics for '2', one emission per matching R row for '3'."""
    r, s = orc.gen_ref(91, 400, num_mod=900)           # 40,000 rows per side, ~44 rows per num: real multiplicities for '3'
    rng = np.random.default_rng(4)
    for img in (r, s):                                  # 7-letter strings from 'a'..'p': 28 varying str bits (+10 of num for '3')
        e = img["entries"].reshape(-1).copy()
        st = np.zeros((len(e), 120), dtype=np.uint8)
        st[:, :7] = rng.integers(ord("a"), ord("q"), size=(len(e), 7), dtype=np.uint8)
        st[::3, 2:] = 0                                 # many short strings: plenty of equal (num, str) pairs
        e["str"] = st.view("V120").reshape(-1)
        img["entries"][:] = e.reshape(img["entries"].shape)
    s["nreserved"][7] = 13                              # a ragged S block
    for field in "23":
        want = orc.hashjoin(r, s, field)
        nres = orc.count_rows(want)
        assert nres > 500
        for no_radix in (False, True):
            if no_radix:
                monkeypatch.setenv("DBT_JOIN_NO_RADIX", "1")
            else:
                monkeypatch.delenv("DBT_JOIN_NO_RADIX", raising=False)
            got, n = H.dev_hashjoin(dbt, orc, r, s, field, cap_blocks=H.nb(nres) + 1)
            assert n == nres and H.same_image(got, want), (field, no_radix, H.first_diff(got, want))
    # 9-letter strings: 36+ varying bits => two key words (hi and lo)
    for img in (r, s):
        e = img["entries"].reshape(-1).copy()
        st = np.zeros((len(e), 120), dtype=np.uint8)
        st[:, :9] = rng.integers(ord("a"), ord("q"), size=(len(e), 9), dtype=np.uint8)
        st[::2, 1:] = 0
        e["str"] = st.view("V120").reshape(-1)
        img["entries"][:] = e.reshape(img["entries"].shape)
    monkeypatch.delenv("DBT_JOIN_NO_RADIX", raising=False)
    for field in "23":
        want = orc.hashjoin(r, s, field)
        got, n = H.dev_hashjoin(dbt, orc, r, s, field, cap_blocks=H.nb(orc.count_rows(want)) + 1)
        assert n == orc.count_rows(want) and H.same_image(got, want), (field, H.first_diff(got, want))
