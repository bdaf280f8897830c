"""Multi-GPU sharding of the tuple operators: one process per GPU, torch.distributed for the plumbing.

SURVEY.md 8e / north_star: sort and dedup shard by sample-sort key-range splitters, joins by key hash,
each with ONE all-to-all over NVLink.  What travels is whole records, as block images: the output of a
distributed sort is range-partitioned across the ranks (rank r holds the r-th key range, so the
concatenation of the ranks' outputs is the globally sorted file), which means (P-1)/P of all records
must cross the fabric whatever the algorithm.

Per rank:   keys = extract(image)                                   [C-ABI, CUDA]
            dest = range bucket by splitters | hash(key) mod P       [C-ABI, CUDA]
            rows grouped by dest (one stable radix pass)             [C-ABI, CUDA]
            P record gathers -> P block images                       [C-ABI, CUDA]
            all_to_all_single of the images                          [NCCL]
            the ordinary single-GPU operator on the received image   [C-ABI, CUDA]

`LocalOps` is the CUDA backend (no fallback).  The orchestration in `DistOps` only needs the small
interface below, which is what lets tests/test_dist_gloo.py drive it on CPU tensors over gloo with a
test-side backend.  All four fields shard: rows are routed on the key's most significant word (equal keys share
it); the columns-only strategies ("keys", "overlap") apply to the u32 fields '0' and '1'.
"""
from __future__ import annotations

import ctypes as C
import importlib

BLOCK_BYTES = 14016
RPB = 100


def _pkg():
    return importlib.import_module("database-technology-algorithms_b200")


class LocalOps:
    """Single-GPU building blocks through the C-ABI of libdbt_b200.so; torch owns the memory."""

    def __init__(self, device):
        import torch

        self.torch = torch
        self.dbt = _pkg()
        self.L = self.dbt.lib()
        if self.L.dbt_device_count() == 0:
            raise RuntimeError("no CUDA device visible: there is no CPU fallback")
        self.device = device
        self._ws = None

    def _stream(self) -> int:
        return self.torch.cuda.current_stream().cuda_stream

    def alloc(self, nbytes: int):
        return self.torch.empty(max(int(nbytes), 256), dtype=self.torch.uint8, device=self.device)

    def workspace(self, nbytes: int):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = self.alloc(nbytes)
        return self._ws

    def extract_keys(self, img, nblocks: int, field: str):
        keys = self.torch.empty(max(nblocks * RPB, 1), dtype=self.torch.int32, device=self.device)
        wsb = self.dbt.dev_ws_bytes(self.dbt.OP_SORT, nblocks, 0, field)
        ws = self.workspace(wsb)
        n = C.c_uint64()
        self.dbt.check(self.L.dbt_dev_extract_keys_u32(img.data_ptr(), nblocks, ord(field), keys.data_ptr(), ws.data_ptr(),
                                                       wsb, self._stream(), C.byref(n)))
        return keys[: n.value]

    def partition(self, keys, mode: int, splitters, nparts: int):
        n = keys.numel()
        rows = self.torch.empty(max(n, 1), dtype=self.torch.int32, device=self.device)
        wsb = self.L.dbt_dev_partition_ws_bytes((n + RPB - 1) // RPB + 1)
        ws = self.workspace(wsb)
        sp = (C.c_uint32 * 64)(*[int(x) for x in splitters]) if splitters else (C.c_uint32 * 64)()
        counts = (C.c_uint64 * 64)()
        self.dbt.check(self.L.dbt_dev_partition_rows(keys.data_ptr(), n, mode, sp, nparts, rows.data_ptr(), counts,
                                                     ws.data_ptr(), wsb, self._stream()))
        return rows[:n], [int(counts[i]) for i in range(nparts)]

    def gather(self, img, rows, out_img):
        self.dbt.check(self.L.dbt_gather_records(img.data_ptr(), rows.data_ptr(), None, rows.numel(), out_img.data_ptr(),
                                                 self._stream()))

    def gather_to_ptr(self, img, rows, out_ptr: int, max_ctas: int = 0):
        """Same kernel, output given as a raw device address (a peer's receive buffer mapped over NVLink)."""
        self.dbt.check(self.L.dbt_gather_records_limited(img.data_ptr(), rows.data_ptr(), None, rows.numel(), out_ptr,
                                                         self._stream(), max_ctas))

    # -- columns for the "rows stay put" sort / dedup ------------------------------------------------
    def extract_key_recid(self, img, nblocks: int, field: str):
        t = self.torch
        keys = t.empty(max(nblocks * RPB, 1), dtype=t.int32, device=self.device)
        recids = t.empty_like(keys)
        wsb = self.dbt.dev_ws_bytes(self.dbt.OP_SORT, nblocks, 0, field)
        ws = self.workspace(wsb)
        n, dense = C.c_uint64(), C.c_int()
        self.dbt.check(self.L.dbt_dev_extract_key_recid_u32(img.data_ptr(), nblocks, ord(field), keys.data_ptr(),
                                                            recids.data_ptr(), ws.data_ptr(), wsb, self._stream(),
                                                            C.byref(n), C.byref(dense)))
        return keys[: n.value], recids[: n.value], bool(dense.value)

    def take(self, src, idx):
        out = self.torch.empty_like(idx)
        self.dbt.check(self.L.dbt_dev_take_u32(src.data_ptr(), idx.data_ptr(), idx.numel(), out.data_ptr(), self._stream()))
        return out

    def order_columns(self, keys, recids, dedup: bool):
        m = keys.numel()
        order = self.torch.empty(max(m, 1), dtype=self.torch.int32, device=self.device)
        wsb = self.L.dbt_dev_order_columns_ws_bytes(m)
        ws = self.workspace(wsb)
        cnt = C.c_uint64()
        self.dbt.check(self.L.dbt_dev_order_columns(keys.data_ptr(), recids.data_ptr(), m, 1 if dedup else 0, order.data_ptr(),
                                                    C.byref(cnt), ws.data_ptr(), wsb, self._stream()))
        return order, cnt.value

    def gather_multi(self, bases, seg_start, order, rrow, count: int, out_img):
        P = len(bases)
        hb = (C.c_void_p * P)(*[C.c_void_p(int(b)) for b in bases])
        hs = (C.c_uint64 * (P + 1))(*[int(x) for x in seg_start])
        self.dbt.check(self.L.dbt_gather_records_multi(hb, P, hs, order.data_ptr(), rrow.data_ptr() if rrow is not None else None,
                                                       count, out_img.data_ptr(), self._stream()))

    def ipc_export(self, tensor):
        handle = C.create_string_buffer(64)
        off = C.c_uint64()
        self.dbt.check(self.L.dbt_ipc_export(tensor.data_ptr(), handle, C.byref(off)))
        return handle.raw, off.value

    # -- peer memory -----------------------------------------------------------------------------
    def ipc_alloc(self, nbytes: int):
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        self.dbt.check(self.L.dbt_ipc_alloc(nbytes, C.byref(ptr), handle))
        return ptr.value, handle.raw

    def ipc_open(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        self.dbt.check(self.L.dbt_ipc_open(C.create_string_buffer(handle, 64), C.byref(ptr)))
        return ptr.value

    def view(self, ptr: int, nbytes: int):
        """torch uint8 view of library-owned device memory (no copy, not owned by torch)."""
        class _Holder:
            __cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}

        return self.torch.as_tensor(_Holder(), device=self.device)

    def sample_keys(self, keys, nsamples: int):
        """Exactly `nsamples` evenly spaced keys as an int64 tensor of unsigned values (-1 = no row)."""
        n = keys.numel()
        if n == 0:
            return self.torch.full((nsamples,), -1, dtype=self.torch.int64, device=self.device)
        idx = (self.torch.arange(nsamples, device=self.device, dtype=self.torch.int64) * n) // nsamples  # exact (float32 linspace is not)
        return keys[idx].to(self.torch.int64) & 0xFFFFFFFF

    def semijoin_keys(self, rkeys, img_s, nb_s: int, field: str):
        """S rows (S order) whose key is in the key column `rkeys` (fields '0'/'1')."""
        dbt = self.dbt
        out = self.alloc(nb_s * BLOCK_BYTES)
        # the workspace bound of a hash join whose build side has as many rows as there are keys
        wsb = dbt.dev_ws_bytes(dbt.OP_HASHJOIN, (rkeys.numel() + RPB - 1) // RPB + 1, nb_s, field)
        n = C.c_uint64()
        dbt.check(self.L.dbt_dev_semijoin_keys(rkeys.data_ptr(), rkeys.numel(), img_s.data_ptr(), nb_s, ord(field), out.data_ptr(),
                                               nb_s, self.workspace(wsb).data_ptr(), wsb, self._stream(), C.byref(n)))
        return out, {"out_rows": n.value}

    def run(self, op: str, field: str, img_r, nb_r: int, img_s=None, nb_s: int = 0):
        """The ordinary device-scope operator on (possibly ragged) images; returns (out image, info)."""
        dbt = self.dbt
        if op == "sort":
            out = self.alloc(nb_r * BLOCK_BYTES)
            wsb = dbt.dev_ws_bytes(dbt.OP_SORT, nb_r, 0, field)
            n = dbt.dev_mergesort(img_r.data_ptr(), nb_r, field, out.data_ptr(), self.workspace(wsb).data_ptr(), wsb, self._stream())
            return out, {"rows": n, "out_rows": n}
        if op == "dedup":
            out = self.alloc(nb_r * BLOCK_BYTES)
            wsb = dbt.dev_ws_bytes(dbt.OP_DEDUP, nb_r, 0, field)
            n, u = dbt.dev_dedup(img_r.data_ptr(), nb_r, field, out.data_ptr(), self.workspace(wsb).data_ptr(), wsb, self._stream())
            return out, {"rows": n, "out_rows": u}
        if op == "hashjoin":
            out = self.alloc(nb_s * BLOCK_BYTES)
            wsb = dbt.dev_ws_bytes(dbt.OP_HASHJOIN, nb_r, nb_s, field)
            k = dbt.dev_hashjoin(img_r.data_ptr(), nb_r, img_s.data_ptr(), nb_s, field, out.data_ptr(), nb_s,
                                 self.workspace(wsb).data_ptr(), wsb, self._stream())
            return out, {"out_rows": k}
        if op == "mergejoin":
            out = self.alloc(min(nb_r, nb_s) * BLOCK_BYTES)
            wsb = dbt.dev_ws_bytes(dbt.OP_MERGEJOIN, nb_r, nb_s, field)
            # no side images here: "1outfile.bin"/"2outfile.bin" belong to the file API (NULL skips their two gathers)
            info = dbt.dev_mergejoin(img_r.data_ptr(), nb_r, img_s.data_ptr(), nb_s, field, None, None,
                                     out.data_ptr(), self.workspace(wsb).data_ptr(), wsb, self._stream())
            return out, {"out_rows": info["nres"], **info}
        raise ValueError(op)


def choose_splitters(all_samples, nparts: int):
    """nparts-1 ascending key values cutting the gathered sample (1-D int64 tensor, -1 = padding) into
    equal parts.  Split on key only, so equal keys always land on the same rank."""
    s = all_samples[all_samples >= 0].sort().values
    if s.numel() == 0 or nparts <= 1:
        return [0] * max(nparts - 1, 0)
    n = s.numel()
    return [int(s[min(n - 1, (n * (i + 1)) // nparts)].item()) for i in range(nparts - 1)]


class DistOps:
    """Sharded operators over a torch.distributed process group (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, ops, group=None, samples_per_rank: int = 16384, peer_exchange=None):
        import os

        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.ops, self.group = torch, dist, ops, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.samples_per_rank = samples_per_rank
        self.last_exchange = {}
        # fused gather + exchange over peer memory (CUDA IPC + NVLink stores); the NCCL all-to-all of
        # gathered images remains for CPU/gloo tests and as an explicit choice (DBT_DIST_EXCHANGE=nccl)
        if peer_exchange is None:
            peer_exchange = hasattr(ops, "ipc_alloc") and os.environ.get("DBT_DIST_EXCHANGE", "p2p") == "p2p"
        self.peer_exchange = bool(peer_exchange) and self.world > 1
        self._recv = {}  # slot -> (capacity bytes, own ptr, [peer ptrs])
        self._opened = {}  # IPC handle bytes -> mapped base pointer
        # sort/dedup strategy on GPUs: "overlap" (default) = key columns decide the order on the main stream
        # while the records are pushed as contiguous block images into the owners' staging buffers on a side
        # stream, then one local gather; "keys" = pull winners from the peers' input images; "records" = exchange
        # whole records first, then the ordinary operator on the received image.
        # measured on B200: pulls win at P=2 (29 vs 32 ms), the push/overlap at P=8 (38 vs 42 ms)
        default_mode = "keys" if self.world <= 2 else "overlap"
        self.sort_mode = os.environ.get("DBT_DIST_SORT", default_mode) if self.peer_exchange else "records"
        self.rows_stay_put = self.sort_mode == "keys"
        self._side = None
        if self.peer_exchange and not self._probe_peer_memory():
            # CUDA IPC / peer access is not usable here (agreed on collectively): the record exchange over NCCL needs neither
            self.peer_exchange = False
            self.sort_mode = "records"
            self.rows_stay_put = False
        self.join_mode = os.environ.get("DBT_DIST_JOIN", "replicate")  # u32 semi-joins: "replicate" R's keys | "partition" both sides
        self.push_ctas = int(os.environ.get("DBT_DIST_PUSH_CTAS", "296"))  # link-bound: 2 CTAs per SM leave room for the main stream

    def _peer_image_bases(self, img):
        """Device pointers to every rank's input image (this rank's own, the others mapped over NVLink)."""
        torch, dist, P = self.torch, self.dist, self.world
        try:
            handle, off = self.ops.ipc_export(img)
            ok = 1
        except Exception:  # noqa: BLE001  (e.g. memory from a VMM allocator cannot be exported)
            handle, off, ok = bytes(64), 0, 0
        rec = torch.tensor(list(handle) + list(int(off).to_bytes(8, "little")) + [ok], dtype=torch.uint8, device=self.ops.device)
        allr = [torch.empty_like(rec) for _ in range(P)]
        dist.all_gather(allr, rec, group=self.group)
        rows = [bytes(t.cpu().tolist()) for t in allr]
        if not all(r[72] for r in rows):
            return None
        bases = []
        for r in range(P):
            if r == self.rank:
                bases.append(img.data_ptr())
                continue
            h, o = rows[r][:64], int.from_bytes(rows[r][64:72], "little")
            if h not in self._opened:
                self._opened[h] = self.ops.ipc_open(h)
            bases.append(self._opened[h] + o)
        return bases

    def _sort_overlap(self, img, nblocks: int, field: str, dedup: bool):
        """Order by exchanging (key, recid) columns (8 B per row) on the main stream; meanwhile a side stream
        pushes the records, grouped by owner and in the same order as the columns, as contiguous block images
        into the owners' staging buffers over NVLink.  The k-th tuple received from rank s is the k-th row of
        s's region of my staging buffer, so the final gather is purely local."""
        torch, dist, P = self.torch, self.dist, self.world
        tl = []

        def mark(name):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            tl.append((name, e))

        mark("start")
        keys, recids, dense = self.ops.extract_key_recid(img, nblocks, field)
        flag = torch.tensor([1 if dense else 0], dtype=torch.int32, device=self.ops.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if not int(flag.item()):
            return None
        mark("extracted")
        splitters = self.splitters_from(self.ops.sample_keys(keys, self.samples_per_rank))
        mark("splitters")
        rows, counts = self.ops.partition(keys, 0, splitters, P)
        mark("partitioned")
        send_blocks = [(c + RPB - 1) // RPB for c in counts]
        info = torch.tensor(counts + send_blocks, dtype=torch.int64, device=self.ops.device)
        allinfo = [torch.empty_like(info) for _ in range(P)]
        dist.all_gather(allinfo, info, group=self.group)  # also orders this step after every rank's previous one
        M = torch.stack(allinfo).cpu().tolist()           # M[src] = counts(dst...) + blocks(dst...)
        rcounts = [int(M[src][self.rank]) for src in range(P)]
        rblocks = [int(M[src][P + self.rank]) for src in range(P)]
        cap, own, peers = self._peer_buffers(0, sum(rblocks) * BLOCK_BYTES)
        # --- main stream first: the small (key, recid) columns cross NVLink alone (8 B per row) ...
        skey, srec = self.ops.take(keys, rows), self.ops.take(recids, rows)
        mark("taken")
        m = sum(rcounts)
        recv = []
        for col in (skey, srec):
            r = torch.empty(max(m, 1), dtype=torch.int32, device=self.ops.device)[:m]
            dist.all_to_all_single(r, col, rcounts, counts, group=self.group)
            recv.append(r)
        mark("columns_exchanged")
        # --- ... then the records are pushed on a side stream while the main stream sorts the columns
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.ops.device)
        main = torch.cuda.current_stream()
        side = self._side
        side.wait_stream(main)
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        with torch.cuda.stream(side):
            ev[0].record()
            row_off = [0] * P
            for d in range(1, P):
                row_off[d] = row_off[d - 1] + counts[d - 1]
            for k in range(P):
                d = (self.rank + 1 + k) % P
                if counts[d] == 0:
                    continue
                blk_off = sum(int(M[src][P + d]) for src in range(self.rank))
                self.ops.gather_to_ptr(img, rows[row_off[d]:row_off[d] + counts[d]], peers[d] + blk_off * BLOCK_BYTES,
                                       max_ctas=self.push_ctas)
            ev[1].record()
        mark("push_enqueued")
        order, cnt = self.ops.order_columns(recv[0], recv[1], dedup)
        mark("ordered")
        main.wait_stream(side)
        dist.barrier(group=self.group)  # every rank's pushes have landed
        mark("pushes_landed")
        out = self.ops.alloc(((cnt + RPB - 1) // RPB) * BLOCK_BYTES)
        seg, bases, boff = [0], [], 0
        for s_ in range(P):
            seg.append(seg[-1] + rcounts[s_])
            bases.append(own + boff * BLOCK_BYTES)
            boff += rblocks[s_]
        self.ops.gather_multi(bases, seg, order, None, cnt, out)
        mark("gathered")
        remote = sum(b for d, b in enumerate(send_blocks) if d != self.rank) * BLOCK_BYTES
        self.last_exchange = {"bytes_sent_remote": remote, "bytes_sent": sum(send_blocks) * BLOCK_BYTES, "events": ev,
                              "timeline": tl,
                              "mode": "columns on main stream || record push on side stream, local final gather",
                              "splitters": splitters}
        return out, {"rows": m, "out_rows": cnt}

    def _sort_rows_stay_put(self, img, nblocks: int, field: str, dedup: bool):
        """Sort / dedup where only (key, recid, row) columns are exchanged (12 B per row) and every rank's
        final gather pulls the records it owns straight out of the peers' input images over NVLink."""
        torch, dist, P = self.torch, self.dist, self.world
        keys, recids, dense = self.ops.extract_key_recid(img, nblocks, field)
        flag = torch.tensor([1 if dense else 0], dtype=torch.int32, device=self.ops.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        bases = self._peer_image_bases(img) if int(flag.item()) else None
        if bases is None:
            return None  # ragged image somewhere, or memory that cannot be mapped: use the record exchange
        splitters = self.splitters_from(self.ops.sample_keys(keys, self.samples_per_rank))
        rows, counts = self.ops.partition(keys, 0, splitters, P)
        skey, srec = self.ops.take(keys, rows), self.ops.take(recids, rows)
        sc = torch.tensor(counts, dtype=torch.int64, device=self.ops.device)
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc, group=self.group)
        rcounts = [int(x) for x in rc.cpu().tolist()]
        m = sum(rcounts)
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        recv = []
        for col in (skey, srec, rows):
            r = torch.empty(max(m, 1), dtype=torch.int32, device=self.ops.device)[:m]
            dist.all_to_all_single(r, col, rcounts, counts, group=self.group)
            recv.append(r)
        ev[1].record()
        rkey, rrec, rrow = recv
        order, cnt = self.ops.order_columns(rkey, rrec, dedup)
        out = self.ops.alloc(((cnt + RPB - 1) // RPB) * BLOCK_BYTES)
        seg = [0]
        for c in rcounts:
            seg.append(seg[-1] + c)
        gev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        gev[0].record()
        self.ops.gather_multi(bases, seg, order, rrow, cnt, out)
        gev[1].record()
        dist.barrier(group=self.group)  # nobody may touch its input image before every peer has read it
        remote_rows = m - rcounts[self.rank]
        self.last_exchange = {"bytes_sent_remote": 12 * (sum(counts) - counts[self.rank]), "bytes_sent": 12 * sum(counts),
                              "events": ev, "gather_events": gev, "mode": "keys+remote-gather",
                              "remote_record_bytes_read": int(remote_rows * (cnt / max(m, 1)) * 140), "splitters": splitters}
        return out, {"rows": m, "out_rows": cnt}

    def _probe_peer_memory(self) -> bool:
        """One-time check, agreed on by all ranks: can every rank map a buffer of every other rank and write to it?"""
        torch, dist, P = self.torch, self.dist, self.world
        ok = 1
        try:
            cap, own, peers = self._peer_buffers(-1, 1 << 20)
            probe = torch.full((64,), self.rank + 1, dtype=torch.uint8, device=self.ops.device)
            dist.barrier(group=self.group)
            for r in range(P):  # every rank writes its id into its own 64-byte slot of everybody's buffer
                self.ops.view(peers[r], 1 << 20)[self.rank * 64:(self.rank + 1) * 64].copy_(probe)
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            mine = self.ops.view(own, 1 << 20)[: P * 64].cpu().view(P, 64)
            ok = int(all(int(mine[r, 0]) == r + 1 for r in range(P)))
        except Exception:  # noqa: BLE001
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.ops.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        return bool(int(flag.item()))

    def _peer_buffers(self, slot: int, need_bytes: int):
        """Receive buffer `slot` of every rank, mapped here; (re)allocated collectively when too small."""
        torch, dist, P = self.torch, self.dist, self.world
        cur = self._recv.get(slot)
        want = torch.tensor([need_bytes], dtype=torch.int64, device=self.ops.device)
        dist.all_reduce(want, op=dist.ReduceOp.MAX, group=self.group)
        need = int(want.item())
        if cur is not None and cur[0] >= need:
            return cur
        cap = need + need // 4 + (1 << 20)
        ptr, handle = self.ops.ipc_alloc(cap)
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.ops.device)
        allh = [torch.empty_like(mine) for _ in range(P)]
        dist.all_gather(allh, mine, group=self.group)
        peers = [ptr if r == self.rank else self.ops.ipc_open(bytes(allh[r].cpu().tolist())) for r in range(P)]
        self._recv[slot] = (cap, ptr, peers)  # (older, smaller buffers are kept alive: peers may still map them)
        return self._recv[slot]

    # -- the one exchange step ---------------------------------------------------------------
    def _exchange_p2p(self, img, rows, counts, send_blocks, splitters, slot):
        """Fused gather + exchange: every per-destination gather writes straight into the destination
        rank's receive buffer (16-byte stores over NVLink); one tiny all-gather of block counts before
        (it also orders this step after every rank's previous consumer) and one barrier after."""
        torch, dist, P = self.torch, self.dist, self.world
        sb = torch.tensor(send_blocks, dtype=torch.int64, device=self.ops.device)
        mat = [torch.empty_like(sb) for _ in range(P)]
        dist.all_gather(mat, sb, group=self.group)
        M = torch.stack(mat).cpu().tolist()  # M[src][dst] blocks
        recv_blocks = [int(M[src][self.rank]) for src in range(P)]
        my_need = sum(recv_blocks) * BLOCK_BYTES
        cap, own, peers = self._peer_buffers(slot, my_need)
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        row_off = [0] * P
        for d in range(1, P):
            row_off[d] = row_off[d - 1] + counts[d - 1]
        for k in range(P):  # rotate the destination order so that the ranks do not all hit the same peer at once
            d = (self.rank + 1 + k) % P
            if counts[d] == 0:
                continue
            blk_off = sum(int(M[src][d]) for src in range(self.rank))
            self.ops.gather_to_ptr(img, rows[row_off[d]:row_off[d] + counts[d]], peers[d] + blk_off * BLOCK_BYTES)
        ev[1].record()
        dist.barrier(group=self.group)  # stream-ordered: every rank's stores have landed before anyone consumes
        remote = sum(b for d, b in enumerate(send_blocks) if d != self.rank) * BLOCK_BYTES
        self.last_exchange = {"bytes_sent_remote": remote, "bytes_sent": sum(send_blocks) * BLOCK_BYTES, "events": ev,
                              "splitters": splitters, "send_rows": counts, "mode": "p2p"}
        return self.ops.view(own, max(my_need, 256)), sum(recv_blocks)

    def exchange(self, img, nblocks: int, field: str, mode: int, splitters=None, slot: int = 0):
        """Route every row of the local image to its owner; returns (received image, nblocks_received).
        The received image is the concatenation, in rank order, of one block image per source rank
        (each with a partial last block at most): a ragged image the device operators accept."""
        torch, dist, P = self.torch, self.dist, self.world
        keys = self.ops.extract_keys(img, nblocks, field)
        if mode == 0 and splitters is None:
            splitters = self.splitters_from(self.ops.sample_keys(keys, self.samples_per_rank))
        rows, counts = self.ops.partition(keys, mode, splitters or [], P)
        send_blocks = [(c + RPB - 1) // RPB for c in counts]
        if self.peer_exchange:
            return self._exchange_p2p(img, rows, counts, send_blocks, splitters, slot)
        send = self.ops.alloc(sum(send_blocks) * BLOCK_BYTES)
        off_rows, off_blocks = 0, 0
        for d in range(P):
            if counts[d]:
                self.ops.gather(img, rows[off_rows:off_rows + counts[d]], send[off_blocks * BLOCK_BYTES:])
            off_rows += counts[d]
            off_blocks += send_blocks[d]
        # sizes: one tiny all-to-all of block counts, then the data
        sb = torch.tensor(send_blocks, dtype=torch.int64, device=send.device)
        rb = torch.empty_like(sb)
        dist.all_to_all_single(rb, sb, group=self.group)
        recv_blocks = [int(x) for x in rb.cpu().tolist()]
        recv = self.ops.alloc(sum(recv_blocks) * BLOCK_BYTES)
        in_split = [b * BLOCK_BYTES for b in send_blocks]
        out_split = [b * BLOCK_BYTES for b in recv_blocks]
        send_v = send[: sum(in_split)]
        recv_v = recv[: sum(out_split)]
        ev = None
        if send.is_cuda:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        dist.all_to_all_single(recv_v, send_v, out_split, in_split, group=self.group)
        if ev:
            ev[1].record()
        self.last_exchange = {"bytes_sent_remote": sum(b for d, b in enumerate(in_split) if d != self.rank),
                              "bytes_sent": sum(in_split), "events": ev, "splitters": splitters, "send_rows": counts,
                              "mode": "nccl"}
        return recv, sum(recv_blocks)

    def splitters_from(self, my_samples):
        gathered = [self.torch.empty_like(my_samples) for _ in range(self.world)]
        self.dist.all_gather(gathered, my_samples, group=self.group)
        return choose_splitters(self.torch.cat(gathered), self.world)

    # -- operators ------------------------------------------------------------------------------
    def sort(self, img, nblocks: int, field: str):
        if self.sort_mode in ("overlap", "keys") and field in ("0", "1"):
            f = self._sort_overlap if self.sort_mode == "overlap" else self._sort_rows_stay_put
            r = f(img, nblocks, field, dedup=False)
            if r is not None:
                return r
        recv, nb = self.exchange(img, nblocks, field, mode=0)
        return self.ops.run("sort", field, recv, nb)

    def dedup(self, img, nblocks: int, field: str):
        if self.sort_mode in ("overlap", "keys") and field in ("0", "1"):
            f = self._sort_overlap if self.sort_mode == "overlap" else self._sort_rows_stay_put
            r = f(img, nblocks, field, dedup=True)
            if r is not None:
                return r
        recv, nb = self.exchange(img, nblocks, field, mode=0)
        return self.ops.run("dedup", field, recv, nb)

    def hashjoin(self, img_r, nb_r: int, img_s, nb_s: int, field: str):
        if field in ("0", "1") and hasattr(self.ops, "semijoin_keys") and self.join_mode == "replicate":
            return self._hashjoin_replicated_keys(img_r, nb_r, img_s, nb_s, field)
        rr, nbr = self.exchange(img_r, nb_r, field, mode=1)
        rs, nbs = self.exchange(img_s, nb_s, field, mode=1, slot=1)
        return self.ops.run("hashjoin", field, rr, nbr, rs, nbs)

    def _hashjoin_replicated_keys(self, img_r, nb_r: int, img_s, nb_s: int, field: str):
        """Semi-join by replicating the build side's KEYS: one all-gather of 4 bytes per R row, then every rank
        probes its own S shard in place.  No S record moves, key skew cannot unbalance the ranks, and the ranks'
        outputs concatenate in S file order (what the reference's HashJoin produces)."""
        torch, dist, P = self.torch, self.dist, self.world
        keys = self.ops.extract_keys(img_r, nb_r, field)
        n_local = torch.tensor([keys.numel()], dtype=torch.int64, device=keys.device)
        sizes = [torch.empty_like(n_local) for _ in range(P)]
        dist.all_gather(sizes, n_local, group=self.group)
        sizes = [int(x.item()) for x in sizes]
        mx = max(max(sizes), 1)
        padded = torch.zeros(mx, dtype=keys.dtype, device=keys.device)
        padded[: keys.numel()] = keys
        allk = torch.empty(mx * P, dtype=keys.dtype, device=keys.device)
        ev = None
        if keys.is_cuda:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        dist.all_gather_into_tensor(allk, padded, group=self.group)
        if ev:
            ev[1].record()
        rkeys = torch.cat([allk[r * mx: r * mx + sizes[r]] for r in range(P)]) if any(s != mx for s in sizes) else allk
        self.last_exchange = {"bytes_sent_remote": 4 * keys.numel() * (P - 1), "bytes_sent": 4 * keys.numel() * P, "events": ev,
                              "mode": "replicated build keys (all-gather), S probed in place"}
        return self.ops.semijoin_keys(rkeys, img_s, nb_s, field)

    def mergejoin(self, img_r, nb_r: int, img_s, nb_s: int, field: str):
        # both relations must use the SAME splitters so that equal keys of R and S meet
        torch, dist, P = self.torch, self.dist, self.world
        kr = self.ops.extract_keys(img_r, nb_r, field)
        ks = self.ops.extract_keys(img_s, nb_s, field)
        sp = self.splitters_from(torch.cat([self.ops.sample_keys(kr, self.samples_per_rank // 2),
                                            self.ops.sample_keys(ks, self.samples_per_rank // 2)]))
        rr, nbr = self.exchange(img_r, nb_r, field, mode=0, splitters=sp)
        rs, nbs = self.exchange(img_s, nb_s, field, mode=0, splitters=sp, slot=1)
        return self.ops.run("mergejoin", field, rr, nbr, rs, nbs)

    def total(self, value: int) -> int:
        t = self.torch.tensor([int(value)], dtype=self.torch.int64,
                              device=self.ops.device if hasattr(self.ops, "device") else "cpu")
        self.dist.all_reduce(t, group=self.group)
        return int(t.item())
