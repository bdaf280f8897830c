"""Independent result checks for bench.py at full scale: plain torch on the device, none of the library's kernels.

The operators only MOVE rows, so a result is right when (1) it holds the right rows -- a 64-bit multiset hash over
all 140 bytes of every record must equal the hash of the input rows that should have been emitted, (2) in the right
order -- sorted by (key, recid) inside a rank and across the rank boundaries, S file order for the hash join -- and
(3) for EliminateDuplicates, every emitted row is THE min-recid row of its key (a direct-address table over the 32-bit
key space built with scatter-min from the inputs of all ranks) and every key is there exactly once.

All functions work on the caller's shard and reduce over `torch.distributed` when `dist` is given; they return a dict of
booleans / numbers (never raise on a wrong result: bench.py reports them and asserts afterwards).
"""
from __future__ import annotations

import torch

BLOCK_WORDS, RPB, REC_WORDS = 3504, 100, 35
_MASK32 = 0xFFFFFFFF


def _words(img_u8, nblocks):
    return img_u8[: nblocks * BLOCK_WORDS * 4].view(torch.int32)


def column(img_u8, nrows: int, word: int):
    """int64 tensor [nrows] of the unsigned 32-bit word `word` (0 recid, 1 num, 2.. str) of every row of a packed image."""
    nb = (nrows + RPB - 1) // RPB
    if nb == 0:
        return torch.zeros(0, dtype=torch.int64, device=img_u8.device)
    t = _words(img_u8, nb)
    v = torch.as_strided(t, (nb, RPB), (BLOCK_WORDS, REC_WORDS), t.storage_offset() + 2 + word)  # (as_strided offsets are absolute)
    return (v.reshape(-1)[:nrows].to(torch.int64)) & _MASK32


def _bswap32(w):
    return ((w & 0xFF) << 24) | ((w & 0xFF00) << 8) | ((w >> 8) & 0xFF00) | ((w >> 24) & 0xFF)


def str_key64(img_u8, nrows: int):
    """The first 8 bytes of str as one big-endian unsigned number (order-preserving for strings shorter than 8 bytes,
    as the generators' 5-letter strings and "Hola" are), shifted down by one bit so that it fits a signed int64."""
    hi, lo = _bswap32(column(img_u8, nrows, 2)), _bswap32(column(img_u8, nrows, 3))
    return ((hi << 32) | lo) >> 1


def composite_key64(img_u8, nrows: int):
    """(num, str) of the generators' rows as one order-preserving int64: num < 2^28, five 7-bit characters."""
    num = column(img_u8, nrows, 1)
    w0, w1 = column(img_u8, nrows, 2), column(img_u8, nrows, 3)
    b = [(w0 >> s) & 0x7F for s in (0, 8, 16, 24)] + [w1 & 0x7F]
    k = num
    for x in b:
        k = (k << 7) | x
    return k


_W = None


def row_hashes(img_u8, nrows: int, chunk_blocks: int = 32768):
    """int64 [nrows]: a mixing hash of all 35 words (140 bytes) of every row (arithmetic wraps modulo 2^64)."""
    global _W
    dev = img_u8.device
    if _W is None or _W.device != dev:
        g = torch.Generator(device="cpu").manual_seed(12345)
        _W = (torch.randint(1, 2**62, (REC_WORDS,), generator=g, dtype=torch.int64) * 2 + 1).to(dev)
    nb = (nrows + RPB - 1) // RPB
    out = torch.empty(nb * RPB, dtype=torch.int64, device=dev)
    t = _words(img_u8, nb)
    for b0 in range(0, nb, chunk_blocks):
        b1 = min(nb, b0 + chunk_blocks)
        v = torch.as_strided(t, (b1 - b0, RPB, REC_WORDS), (BLOCK_WORDS, REC_WORDS, 1), t.storage_offset() + 2 + b0 * BLOCK_WORDS)
        h = ((v.to(torch.int64) & _MASK32) * _W).sum(-1)
        h = (h ^ (h >> 29)) * -4658895280553007687  # 0xBF58476D1CE4E5B9 as int64
        h = h ^ (h >> 32)
        out[b0 * RPB: b1 * RPB] = h.reshape(-1)
    return out[:nrows]


def _allsum(x: int, dist, dev):
    t = torch.tensor([x], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(t)
    return int(t.item())


def _boundaries_ok(first_key, first_rid, last_key, last_rid, nrows, strict: bool, dist, dev):
    """Global order across the rank boundaries: rank r's last (key, recid) precedes rank r+1's first."""
    if dist is None:
        return True
    mine = torch.tensor([first_key, first_rid, last_key, last_rid, nrows], dtype=torch.int64, device=dev)
    every = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(every, mine)
    rows = [e.tolist() for e in every if int(e[4]) > 0]
    ok = True
    for a, b in zip(rows, rows[1:]):
        ok = ok and ((a[2] < b[0]) or (not strict and a[2] == b[0] and a[3] < b[1]))
    return ok


def check_sort(in_img, n_in: int, out_img, n_out: int, key_fn, dist=None):
    """MergeSort: every input row exactly once, ordered by (key, recid) inside the rank and across ranks."""
    dev = in_img.device
    k, r = key_fn(out_img, n_out), column(out_img, n_out, 0)
    ordered = True
    if n_out > 1:
        ordered = bool(((k[1:] > k[:-1]) | ((k[1:] == k[:-1]) & (r[1:] > r[:-1]))).all().item())
    b_ok = _boundaries_ok(int(k[0]) if n_out else 0, int(r[0]) if n_out else 0, int(k[-1]) if n_out else 0,
                          int(r[-1]) if n_out else 0, n_out, False, dist, dev)
    h_in = _allsum(int(row_hashes(in_img, n_in).sum().item()), dist, dev)
    h_out = _allsum(int(row_hashes(out_img, n_out).sum().item()), dist, dev)
    rows_in, rows_out = _allsum(n_in, dist, dev), _allsum(n_out, dist, dev)
    return {"ordered_in_rank": ordered, "ordered_across_ranks": b_ok, "rows": rows_out, "rows_expected": rows_in,
            "record_multiset_hash_equal": h_in == h_out and rows_in == rows_out}


def check_dedup_u32(in_img, n_in: int, out_img, n_out: int, word: int, dist=None, key_space: int = 2**32):
    """EliminateDuplicates on a 32-bit key (word 0 recid / 1 num): strictly ascending keys, every emitted row is the
    min-recid row of its key, every key exactly once, emitted records byte-identical to the winning input rows."""
    dev = in_img.device
    sent = 2**31 - 1
    k_in, r_in = column(in_img, n_in, word), column(in_img, n_in, 0)
    table = torch.full((key_space,), sent, dtype=torch.int32, device=dev)  # 16 GB for 32-bit keys: min recid per possible key
    table.scatter_reduce_(0, k_in, r_in.to(torch.int32), reduce="amin")
    if dist is not None:
        dist.all_reduce(table, op=dist.ReduceOp.MIN)
    k, r = column(out_img, n_out, word), column(out_img, n_out, 0)
    strictly = bool((k[1:] > k[:-1]).all().item()) if n_out > 1 else True
    is_min = bool((table[k].to(torch.int64) == r).all().item()) if n_out else True
    b_ok = _boundaries_ok(int(k[0]) if n_out else 0, 0, int(k[-1]) if n_out else 0, 0, n_out, True, dist, dev)
    distinct = int((table != sent).sum().item())
    winners = table[k_in].to(torch.int64) == r_in
    h_in = _allsum(int(row_hashes(in_img, n_in)[winners].sum().item()), dist, dev)
    h_out = _allsum(int(row_hashes(out_img, n_out).sum().item()), dist, dev)
    rows_out = _allsum(n_out, dist, dev)
    del table
    return {"keys_strictly_ascending": strictly, "ordered_across_ranks": b_ok, "every_row_is_min_recid_of_its_key": is_min,
            "rows": rows_out, "distinct_keys_in_input": distinct, "record_multiset_hash_equal": h_in == h_out and rows_out == distinct}


def check_semijoin_u32(r_img, n_r: int, s_img, n_s: int, out_img, n_out: int, word: int, domain: int, dist=None,
                       chunk_rows: int = 50_000_000):
    """HashJoin fields '0'/'1' (set semantics): the output is exactly the S rows, in S file order, whose key is in
    keys(R) -- compared row by row, all 140 bytes, against a torch boolean-table filter of the S shard (S is walked
    in chunks of whole blocks so that the temporaries stay small)."""
    dev = s_img.device
    table = torch.zeros(domain, dtype=torch.uint8, device=dev)
    kr = column(r_img, n_r, word)
    table[kr[kr < domain]] = 1
    if dist is not None:
        dist.all_reduce(table, op=dist.ReduceOp.MAX)
    del kr
    chunk_rows = max(RPB, chunk_rows // RPB * RPB)
    want, same_rows, pos = 0, True, 0
    for r0 in range(0, n_s, chunk_rows):
        m = min(chunk_rows, n_s - r0)
        sub = s_img[(r0 // RPB) * BLOCK_WORDS * 4:]
        ks = column(sub, m, word)
        inside = ks < domain
        mask = torch.zeros(m, dtype=torch.bool, device=dev)
        mask[inside] = table[ks[inside]].bool()
        k = int(mask.sum().item())
        want += k
        if same_rows and pos + k <= n_out and k:
            hs = row_hashes(sub, m)[mask]
            # output rows [pos, pos + k): they start in the middle of an output block, so hash a block-aligned window
            b0 = pos // RPB
            ho = row_hashes(out_img[b0 * BLOCK_WORDS * 4:], (pos - b0 * RPB) + k)[pos - b0 * RPB:]
            same_rows = bool((hs == ho).all().item())  # position by position: same rows in the same (S file) order
            del hs, ho
        elif pos + k > n_out:
            same_rows = False
        pos += k
        del ks, inside, mask
    del table
    return {"rows": _allsum(n_out, dist, dev), "rows_expected": _allsum(want, dist, dev), "same_rows_in_s_order": same_rows and want == n_out}


def check_mergejoin_composite(r_img, n_r: int, s_img, n_s: int, out_img, n_out: int, dist=None):
    """MergeJoin field '3' on generator rows: R's min-recid row of every (num, str) key present in both relations,
    ascending key order across the ranks.  The expectation is computed from the gathered key columns with torch.sort."""
    dev = r_img.device

    def gathered(keys, recids):
        if dist is None:
            return keys, recids
        n = torch.tensor([keys.numel()], dtype=torch.int64, device=dev)
        sizes = [torch.empty_like(n) for _ in range(dist.get_world_size())]
        dist.all_gather(sizes, n)
        sizes = [int(x.item()) for x in sizes]
        mx = max(sizes)
        outk, outr = [], []
        for col, acc in ((keys, outk), (recids, outr)):
            pad = torch.zeros(mx, dtype=torch.int64, device=dev)
            pad[: col.numel()] = col
            every = [torch.empty_like(pad) for _ in sizes]
            dist.all_gather(every, pad)
            acc.extend(e[:m] for e, m in zip(every, sizes))
        return torch.cat(outk), torch.cat(outr)

    kr, rr = gathered(composite_key64(r_img, n_r), column(r_img, n_r, 0))
    ks, _ = gathered(composite_key64(s_img, n_s), column(s_img, n_s, 0))
    order = torch.argsort(kr, stable=True)  # ranks hold ascending recid ranges: stable => min recid first
    kr_s, rr_s = kr[order], rr[order]
    first = torch.ones_like(kr_s, dtype=torch.bool)
    first[1:] = kr_s[1:] != kr_s[:-1]
    uk, ur = kr_s[first], rr_s[first]
    del order, kr_s, rr_s, first
    us = torch.unique(ks)
    hit = torch.isin(uk, us)
    want_k, want_r = uk[hit], ur[hit]
    total = int(want_k.numel())
    mine = torch.tensor([n_out], dtype=torch.int64, device=dev)
    off = 0
    if dist is not None:
        every = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(every, mine)
        off = sum(int(e.item()) for e in every[: dist.get_rank()])
    ko, ro = composite_key64(out_img, n_out), column(out_img, n_out, 0)
    same = off + n_out <= total and bool(((want_k[off: off + n_out] == ko) & (want_r[off: off + n_out] == ro)).all().item())
    rows = _allsum(n_out, dist, dev)
    # content: the emitted records are byte-identical to R's winning rows (wherever those live)
    local_win = torch.isin(column(r_img, n_r, 0), want_r)
    h_in = _allsum(int(row_hashes(r_img, n_r)[local_win].sum().item()), dist, dev)
    h_out = _allsum(int(row_hashes(out_img, n_out).sum().item()), dist, dev)
    return {"rows": rows, "rows_expected": total, "same_keys_and_recids_in_global_order": same and rows == total,
            "record_multiset_hash_equal": h_in == h_out}
