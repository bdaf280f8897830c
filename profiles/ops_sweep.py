"""Device-scope timing of every operator x field on reference-generator-like data (G_syn kind 1).
Usage: python profiles/ops_sweep.py <rows> [ops] [fields]"""
import ctypes as C, importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dbt = importlib.import_module("database-technology-algorithms_b200"); L = dbt.lib()
n = int(sys.argv[1]); ops = (sys.argv[2] if len(sys.argv) > 2 else "sort,dedup,mergejoin,hashjoin").split(",")
fields = sys.argv[3] if len(sys.argv) > 3 else "0123"
BB = 14016; nb = n // 100; sp = torch.cuda.current_stream().cuda_stream
U = n * 3 // 10  # ~3.3 rows per key, like main.cpp (num = rand() % (nblocks*30))
d_r = torch.empty(nb * BB, dtype=torch.uint8, device="cuda"); d_s = torch.empty(nb * BB, dtype=torch.uint8, device="cuda")
dbt.check(L.dbt_gen_syn(7, n, U, 1, 0, n, 0, d_r.data_ptr(), sp)); dbt.check(L.dbt_gen_syn(9, n, U, 1, 0, n, 0, d_s.data_ptr(), sp))
o1, o2, o3 = (torch.empty(nb * BB, dtype=torch.uint8, device="cuda") for _ in range(3))
L.dbt_stage_timing_enable(1)
for field in fields:
    for op in ops:
        opid = {"sort": dbt.OP_SORT, "dedup": dbt.OP_DEDUP, "mergejoin": dbt.OP_MERGEJOIN, "hashjoin": dbt.OP_HASHJOIN}[op]
        wsb = dbt.dev_ws_bytes(opid, nb, nb if "join" in op else 0, field); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
        ts = []; res = None
        for it in range(4):
            torch.cuda.synchronize(); L.dbt_stage_timing_reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
            if op == "sort": res = dbt.dev_mergesort(d_r.data_ptr(), nb, field, o1.data_ptr(), ws.data_ptr(), wsb, sp)
            elif op == "dedup": res = dbt.dev_dedup(d_r.data_ptr(), nb, field, o1.data_ptr(), ws.data_ptr(), wsb, sp)
            elif op == "mergejoin": res = dbt.dev_mergejoin(d_r.data_ptr(), nb, d_s.data_ptr(), nb, field, o1.data_ptr(), o2.data_ptr(), o3.data_ptr(), ws.data_ptr(), wsb, sp)["nres"]
            else: res = dbt.dev_hashjoin(d_r.data_ptr(), nb, d_s.data_ptr(), nb, field, o1.data_ptr(), nb, ws.data_ptr(), wsb, sp)
            e1.record(); torch.cuda.synchronize()
            if it >= 1: ts.append(e0.elapsed_time(e1))
        rep = dbt.stage_report(); ms = sum(ts) / len(ts)
        print(f"field {field} {op:9s} n={n}: {ms:8.2f} ms  {n/ms/1e6:6.2f} G rows/s  result={res}  " + " ".join(f"{a}={b[0]:.2f}" for a, b in rep.items()), flush=True)
        del ws
