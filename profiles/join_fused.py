#!/usr/bin/env python
"""HashJoin field=num, R=100M x S=<rows> at device scope: the fused streaming semi-join against the column-based path
(DBT_JOIN_FUSED=0), same inputs, same box.  usage: python profiles/join_fused.py [s_rows] [ctas_per_sm...]"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dbt = importlib.import_module("database-technology-algorithms_b200")
L = dbt.lib()
RPB, BB = 100, 14016
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000_000
nr, D = 100_000_000, 100_000_000
dev = torch.device("cuda", 0)
sp = torch.cuda.current_stream().cuda_stream
nbr, nbs = nr // RPB, ns // RPB
d_r = torch.empty(nbr * BB, dtype=torch.uint8, device=dev)
d_s = torch.empty(nbs * BB, dtype=torch.uint8, device=dev)
d_o = torch.empty(nbs * BB, dtype=torch.uint8, device=dev)
dbt.check(L.dbt_gen_syn(7, nr, D, 1, 0, nr, 0, d_r.data_ptr(), sp))
wsb = dbt.dev_ws_bytes(dbt.OP_HASHJOIN, nbr, nbs, "1")
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
peak = 6539.9
for kind, label in ((1, "uniform"), (4, "zipf1.1")):
    dbt.check(L.dbt_gen_syn(9, ns, D, kind, 0, ns, 0, d_s.data_ptr(), sp))
    ref = None
    for fused in os.environ.get("JOIN_MODES", "2,1,0").split(","):
        os.environ["DBT_JOIN_FUSED"] = fused
        L.dbt_stage_timing_enable(1)
        times = []
        for it in range(4):
            torch.cuda.synchronize()
            L.dbt_stage_timing_reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            k = dbt.dev_hashjoin(d_r.data_ptr(), nbr, d_s.data_ptr(), nbs, "1", d_o.data_ptr(), nbs, ws.data_ptr(), wsb, sp)
            e1.record()
            torch.cuda.synchronize()
            if it:
                times.append(e0.elapsed_time(e1))
        rep = dbt.stage_report()
        ms = sum(times) / len(times)
        nbo = (k + 99) // 100
        chk = int(d_o[: nbo * BB].view(torch.int32)[:: 997].to(torch.int64).sum().item())  # sampled checksum of the output image
        if ref is None:
            ref = (k, chk)
        probe_ms = rep["hash_probe"][0] + (rep.get("record_gather", (0, 0))[0] + rep.get("compact", (0, 0))[0] if fused == "2" else 0)
        alg = (140.0 * ns + 140.0 * k) if fused != "0" else None
        print(json.dumps({"S_rows": ns, "dist": label, "fused": fused, "ms": round(ms, 3), "nres": k, "same_as_fused": (k, chk) == ref,
                          "stage_ms": {a: round(b[0], 3) for a, b in rep.items()},
                          "stream_pass_hbm_frac": round(alg / (probe_ms * 1e-3) / 1e9 / peak, 4) if alg else None}), flush=True)
