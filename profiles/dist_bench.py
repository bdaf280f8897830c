#!/usr/bin/env python
"""Under torchrun: distributed EliminateDuplicates / MergeSort through the C++ layer (dbt.Dist), rows per GPU fixed,
for several pipeline depths (key sub-ranges per owner).  usage: torchrun ... profiles/dist_bench.py [rows_per_gpu] [q,q,...] [field] [dedup 0|1]"""
import importlib
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dbt = importlib.import_module("database-technology-algorithms_b200")
L = dbt.lib()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
qs = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,4,8").split(",")]
field = sys.argv[3] if len(sys.argv) > 3 else "1"
dedup = (sys.argv[4] if len(sys.argv) > 4 else "1") == "1"
tok = [f"{os.environ.get('MASTER_PORT', '0')}_{os.getpid()}_{int(time.time() * 1e3) % 100000}"]
dist.broadcast_object_list(tok, src=0)
d = dbt.Dist(tok[0], rank, world, lr)
RPB, BB = 100, 14016
nb = n // RPB
n_total = n * world
U = n_total * 9 // 10
sp = torch.cuda.current_stream().cuda_stream
d_in = torch.empty(nb * BB, dtype=torch.uint8, device=dev)
dbt.check(L.dbt_gen_syn(42, n_total, U, 0, rank * n, n, 0, d_in.data_ptr(), sp))
cap = int(nb * 1.25) + 64
d_out = torch.empty(cap * BB, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
STAGES = os.environ.get("DIST_BENCH_STAGES") == "1"
if STAGES:
    L.dbt_stage_timing_enable(1)
for q in qs:
    d.set_sub_ranges(q)
    times = []
    for it in range(5):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if STAGES and it == 4:
            L.dbt_stage_timing_reset()
        e0.record()
        rows, recv = d.sort(d_in.data_ptr(), nb, field, dedup, d_out.data_ptr(), cap, sp)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            times.append(e0.elapsed_time(e1))
    st = d.stats()
    ms = torch.tensor([sum(times) / len(times)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    tot = torch.tensor([rows], device=dev, dtype=torch.int64)
    dist.all_reduce(tot)
    if rank == 0:
        gbs = st["bytes_remote"] / (st["nvlink_ms"] * 1e-3) / 1e9 if st["nvlink_ms"] else None
        print(json.dumps({"world": world, "rows_per_gpu": n, "field": field, "dedup": dedup, "sub_ranges": st["sub_ranges"], "ms": round(float(ms.item()), 3),
                          "records_per_s": n_total / (float(ms.item()) * 1e-3), "out_rows_total": int(tot.item()), "expected": U if dedup else n_total,
                          "push_ms_rank0": round(st["nvlink_ms"], 3), "nvlink_gbs_per_direction_rank0": gbs,
                          "timeline_ms_rank0": st["timeline_ms"]}), flush=True)
        if STAGES:
            print(json.dumps({"stages_last_iteration_rank0": dbt.stage_report()}), flush=True)
d.barrier()
d.close()
dist.barrier()
dist.destroy_process_group()
