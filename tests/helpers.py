"""Test-side plumbing: torch owns device memory, the product is called through its C-ABI."""
from __future__ import annotations

import numpy as np

BLOCK_BYTES = 14016
RPB = 100


def _torch():
    import torch

    return torch


def to_dev(blocks: np.ndarray):
    torch = _torch()
    raw = np.ascontiguousarray(blocks).view(np.uint8).reshape(-1)
    t = torch.empty(max(raw.size, 16), dtype=torch.uint8, device="cuda")
    if raw.size:
        t[: raw.size].copy_(torch.from_numpy(raw.copy()))
    return t


def dev_alloc(nbytes: int):
    return _torch().empty(max(int(nbytes), 256), dtype=_torch().uint8, device="cuda")


def to_host(t, nblocks: int, orc) -> np.ndarray:
    raw = t[: nblocks * BLOCK_BYTES].cpu().numpy()
    return orc.as_blocks(raw.copy())


def nb(rows: int) -> int:
    return (rows + RPB - 1) // RPB


def stream() -> int:
    return _torch().cuda.current_stream().cuda_stream


def dev_sort(dbt, orc, blocks, field, kw=8):
    d_in = to_dev(blocks)
    d_out = dev_alloc(len(blocks) * BLOCK_BYTES)
    wsb = dbt.dev_ws_bytes(dbt.OP_SORT, len(blocks), 0, field, kw)
    ws = dev_alloc(wsb)
    n = dbt.dev_mergesort(d_in.data_ptr(), len(blocks), field, d_out.data_ptr(), ws.data_ptr(), wsb, stream())
    return to_host(d_out, nb(n), orc), n


def dev_dedup(dbt, orc, blocks, field, kw=8):
    d_in = to_dev(blocks)
    d_out = dev_alloc(len(blocks) * BLOCK_BYTES)
    wsb = dbt.dev_ws_bytes(dbt.OP_DEDUP, len(blocks), 0, field, kw)
    ws = dev_alloc(wsb)
    n, u = dbt.dev_dedup(d_in.data_ptr(), len(blocks), field, d_out.data_ptr(), ws.data_ptr(), wsb, stream())
    return to_host(d_out, nb(u), orc), n, u


def dev_hashjoin(dbt, orc, r, s, field, cap_blocks=None, kw=8):
    d_r, d_s = to_dev(r), to_dev(s)
    cap = len(s) if cap_blocks is None else cap_blocks
    d_out = dev_alloc(cap * BLOCK_BYTES)
    wsb = dbt.dev_ws_bytes(dbt.OP_HASHJOIN, len(r), len(s), field, kw)
    ws = dev_alloc(wsb)
    n = dbt.dev_hashjoin(d_r.data_ptr(), len(r), d_s.data_ptr(), len(s), field, d_out.data_ptr(), cap, ws.data_ptr(),
                         wsb, stream())
    return to_host(d_out, nb(n), orc), n


def dev_mergejoin(dbt, orc, r, s, field, kw=8):
    d_r, d_s = to_dev(r), to_dev(s)
    d_ur, d_us = dev_alloc(len(r) * BLOCK_BYTES), dev_alloc(len(s) * BLOCK_BYTES)
    d_out = dev_alloc(min(len(r), len(s)) * BLOCK_BYTES)
    wsb = dbt.dev_ws_bytes(dbt.OP_MERGEJOIN, len(r), len(s), field, kw)
    ws = dev_alloc(wsb)
    info = dbt.dev_mergejoin(d_r.data_ptr(), len(r), d_s.data_ptr(), len(s), field, d_ur.data_ptr(), d_us.data_ptr(),
                             d_out.data_ptr(), ws.data_ptr(), wsb, stream())
    return (to_host(d_out, nb(info["nres"]), orc), to_host(d_ur, nb(info["nunique_R"]), orc),
            to_host(d_us, nb(info["nunique_S"]), orc), info)


def same_image(a: np.ndarray, b: np.ndarray) -> bool:
    return a.shape == b.shape and a.tobytes() == b.tobytes()


def first_diff(a: np.ndarray, b: np.ndarray) -> str:
    if a.shape != b.shape:
        return f"block counts differ: {a.shape} vs {b.shape}"
    x = a.view(np.uint8).reshape(len(a), -1)
    y = b.view(np.uint8).reshape(len(b), -1)
    bad = np.argwhere(x != y)
    if bad.size == 0:
        return "identical"
    blk, off = bad[0]
    return f"{len(bad)} differing bytes; first at block {blk} byte {off} ({x[blk, off]} vs {y[blk, off]})"
