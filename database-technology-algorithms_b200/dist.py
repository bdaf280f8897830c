"""Algorithm harness of the multi-GPU sharding, for the CPU (gloo) tests only.

The PRODUCT multi-GPU path is C++: csrc/dist.cu (`dbt_dist_*` of include/dbt_b200.h, `Dist` in this package) --
shared-memory control block, per-owner block images moved over NVLink by the copy engines, pipelined key sub-ranges.
It cannot run without GPUs, so the sharding ALGORITHM it implements is restated here over `torch.distributed` with a
pluggable per-rank backend: tests/test_dist_gloo.py drives it on CPU tensors over gloo (world sizes 2 and 3) with an
oracle-backed backend and checks that it composes to exactly the single-node CANON result:

  sort / dedup / mergejoin   key-range partition with sample-sort splitters (split on the key only, so equal keys meet;
                             the ranks' outputs concatenate to the globally ordered file), mergejoin with the SAME
                             splitters for both relations;
  hashjoin, fields '0'/'1'   the build side's KEYS are replicated and every rank probes its own S shard in place (no S
                             record moves, the outputs concatenate in S file order);
  hashjoin, fields '2'/'3'   both relations hash-partitioned on the key's first word.

Rows are routed on the key's most significant word (recid, num, or the first four str bytes): equal keys share it.
"""
from __future__ import annotations

BLOCK_BYTES = 14016
RPB = 100


def choose_splitters(all_samples, nparts: int):
    """nparts-1 ascending key values cutting the gathered sample (1-D int64 tensor, -1 = padding) into
    equal parts.  Split on key only, so equal keys always land on the same rank."""
    s = all_samples[all_samples >= 0].sort().values
    if s.numel() == 0 or nparts <= 1:
        return [0] * max(nparts - 1, 0)
    n = s.numel()
    return [int(s[min(n - 1, (n * (i + 1)) // nparts)].item()) for i in range(nparts - 1)]


class DistOps:
    """The sharding algorithm over a torch.distributed process group (gloo in the CPU tests); `ops` is the per-rank backend."""

    def __init__(self, ops, group=None, samples_per_rank: int = 16384):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.ops, self.group = torch, dist, ops, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.samples_per_rank = samples_per_rank
        self.last_exchange = {}

    # -- the one exchange step ---------------------------------------------------------------
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
        recv, nb = self.exchange(img, nblocks, field, mode=0)
        return self.ops.run("sort", field, recv, nb)

    def dedup(self, img, nblocks: int, field: str):
        recv, nb = self.exchange(img, nblocks, field, mode=0)
        return self.ops.run("dedup", field, recv, nb)

    def hashjoin(self, img_r, nb_r: int, img_s, nb_s: int, field: str):
        if field in ("0", "1") and hasattr(self.ops, "semijoin_keys"):
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
