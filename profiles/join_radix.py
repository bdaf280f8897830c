#!/usr/bin/env python
"""HashJoin fields '3' and '2' (reference-like composite / string keys), device scope: radix-partitioned build/probe with
per-partition shared-memory tables against the linear-probing table in HBM (DBT_JOIN_NO_RADIX=1).  usage: [rows per side]"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
dbt = importlib.import_module("database-technology-algorithms_b200"); L = dbt.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
BB = 14016; nb = n // 100
sp = torch.cuda.current_stream().cuda_stream
d_r = torch.empty(nb * BB, dtype=torch.uint8, device="cuda"); d_s = torch.empty(nb * BB, dtype=torch.uint8, device="cuda")
dbt.check(L.dbt_gen_syn(21, n, int(0.3 * n), 1, 0, n, 0, d_r.data_ptr(), sp))
dbt.check(L.dbt_gen_syn(21 ^ 0x5EED, n, int(0.3 * n), 3, 0, n, 0, d_s.data_ptr(), sp))  # about half of S's (num, str) keys exist in R
for field in "32":
    cap = nb * (3 if field == "3" else 1)
    d_o = torch.empty(cap * BB, dtype=torch.uint8, device="cuda")
    wsb = int(L.dbt_dev_hashjoin_ws_bytes(nb, nb, ord(field), 8, cap)); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    ref = None
    for no_radix in ("", "1"):
        if no_radix: os.environ["DBT_JOIN_NO_RADIX"] = "1"
        else: os.environ.pop("DBT_JOIN_NO_RADIX", None)
        L.dbt_stage_timing_enable(1); ts = []
        for it in range(4):
            torch.cuda.synchronize(); L.dbt_stage_timing_reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
            k = dbt.dev_hashjoin(d_r.data_ptr(), nb, d_s.data_ptr(), nb, field, d_o.data_ptr(), cap, ws.data_ptr(), wsb, sp)
            e1.record(); torch.cuda.synchronize()
            if it: ts.append(e0.elapsed_time(e1))
        rep = dbt.stage_report()
        chk = int(d_o[: ((k + 99) // 100) * BB].view(torch.int32)[::997].to(torch.int64).sum().item())
        ref = ref or (k, chk)
        print(json.dumps({"rows_per_side": n, "field": field, "path": "hbm table" if no_radix else "radix + smem tables", "ms": round(sum(ts) / len(ts), 3),
                          "nres": k, "same_output": (k, chk) == ref, "stage_ms": {a: round(b[0], 3) for a, b in rep.items()}}), flush=True)
    del d_o, ws
