#!/usr/bin/env python
"""Where an e2e step goes: the headline workload (EliminateDuplicates field=num, 100M records, pinned host image in / out)
through the blocking call and through two job slots, with the C-ABI's stage timing on (H2D / D2H copy durations as CUDA
events on each slot's stream) and plain pinned copies of the same sizes beside them.
usage: python profiles/e2e_timeline.py [rows]"""
import ctypes as C
import importlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dbt = importlib.import_module("database-technology-algorithms_b200")
L = dbt.lib()
RPB, BB = 100, 14016
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
U = n * 9 // 10
nb = n // RPB
img = nb * BB
outb = ((U + RPB - 1) // RPB) * BB
dev = torch.device("cuda", 0)
sp = torch.cuda.current_stream().cuda_stream
d_in = torch.empty(img, dtype=torch.uint8, device=dev)
dbt.check(L.dbt_gen_syn(42, n, U, 0, 0, n, 0, d_in.data_ptr(), sp))
h_in, h_o1, h_o2 = C.c_void_p(), C.c_void_p(), C.c_void_p()
for h, b in ((h_in, img), (h_o1, img), (h_o2, outb)):
    dbt.check(L.dbt_host_alloc(C.byref(h), b))
host_t = torch.frombuffer((C.c_uint8 * img).from_address(h_in.value), dtype=torch.uint8)
host_t.copy_(d_in)
torch.cuda.synchronize()


def plain_copies():
    """pinned cudaMemcpyAsync of the same sizes: up alone, down alone, both at once (two streams)"""
    o1 = torch.frombuffer((C.c_uint8 * outb).from_address(h_o1.value), dtype=torch.uint8)
    d_o = torch.empty(outb, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for name, up, down in (("up_alone", True, False), ("down_alone", False, True), ("both", True, True)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        if up:
            with torch.cuda.stream(s1):
                e[0].record()
                d_in.copy_(host_t, non_blocking=True)
                e[1].record()
        if down:
            with torch.cuda.stream(s2):
                e[2].record()
                o1.copy_(d_o, non_blocking=True)
                e[3].record()
        torch.cuda.synchronize()
        res[name] = {"wall_ms": round((time.perf_counter() - t0) * 1e3, 1)}
        if up:
            res[name]["up_ms"] = round(e[0].elapsed_time(e[1]), 1)
            res[name]["up_GBs"] = round(img / e[0].elapsed_time(e[1]) / 1e6, 1)
        if down:
            res[name]["down_ms"] = round(e[2].elapsed_time(e[3]), 1)
            res[name]["down_GBs"] = round(outb / e[2].elapsed_time(e[3]) / 1e6, 1)
    del d_o
    return res


print(json.dumps({"plain_pinned_copies": plain_copies()}), flush=True)
del d_in
torch.cuda.empty_cache()
nr, nu = C.c_uint64(), C.c_uint64()
L.dbt_stage_timing_enable(1)
for it in range(3):
    L.dbt_stage_timing_reset()
    t0 = time.perf_counter()
    dbt.check(L.dbt_host_dedup(h_in, nb, ord("1"), h_o1, 0, C.byref(nr), C.byref(nu)))
    ms = (time.perf_counter() - t0) * 1e3
    if it:
        print(json.dumps({"blocking_call_ms": round(ms, 1), "stage_ms": {k: round(v[0], 2) for k, v in dbt.stage_report().items()}}), flush=True)
res4 = (C.c_uint64 * 4)()
outs = [h_o1, h_o2]


def pipelined(steps, trace=None):
    pending = [False, False]
    t00 = time.perf_counter()
    for i in range(steps):
        sl = i % 2
        if pending[sl]:
            dbt.check(L.dbt_host_job_wait(sl, res4))
            if trace is not None:
                trace.append(("wait_done", i - 2, round((time.perf_counter() - t00) * 1e3, 1)))
        dbt.check(L.dbt_host_dedup_begin(sl, h_in, nb, ord("1"), outs[sl], 0))
        if trace is not None:
            trace.append(("begin_returned", i, round((time.perf_counter() - t00) * 1e3, 1)))
        pending[sl] = True
    for sl in ((steps % 2), 1 - (steps % 2)):
        if pending[sl]:
            dbt.check(L.dbt_host_job_wait(sl, res4))
            if trace is not None:
                trace.append(("wait_done", "tail", round((time.perf_counter() - t00) * 1e3, 1)))
    return (time.perf_counter() - t00) * 1e3


L.dbt_stage_timing_enable(0)  # stage timing resolves its events with host syncs: two slots would no longer overlap
pipelined(3)
tr = []
ms = pipelined(6, tr)
print(json.dumps({"pipelined_6_steps_ms": round(ms, 1), "per_step": round(ms / 6, 1), "host_trace_ms": tr}), flush=True)
