"""Times the device-scope HashJoin (field num): R and S images generated in HBM.
Usage: python profiles/join_sweep.py <rows_R> <rows_S> [kind_S: 1 uniform | 2 skewed] [domain]"""
import ctypes as C, importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dbt = importlib.import_module("database-technology-algorithms_b200"); L = dbt.lib()
nr, ns = int(sys.argv[1]), int(sys.argv[2]); kind = int(sys.argv[3]) if len(sys.argv) > 3 else 1
D = int(sys.argv[4]) if len(sys.argv) > 4 else nr
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
BB = 14016; nbr, nbs = nr // 100, ns // 100
sp = torch.cuda.current_stream().cuda_stream
d_r = torch.empty(nbr * BB, dtype=torch.uint8, device="cuda"); d_s = torch.empty(nbs * BB, dtype=torch.uint8, device="cuda")
d_o = torch.empty(nbs * BB, dtype=torch.uint8, device="cuda")
dbt.check(L.dbt_gen_syn(7, nr, D, 1, 0, nr, 0, d_r.data_ptr(), sp))
dbt.check(L.dbt_gen_syn(9, ns, D, kind, 0, ns, 0, d_s.data_ptr(), sp))
wsb = dbt.dev_ws_bytes(dbt.OP_HASHJOIN, nbr, nbs, "1"); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
L.dbt_stage_timing_enable(1)
ts = []
for it in range(5):
    torch.cuda.synchronize(); L.dbt_stage_timing_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    k = dbt.dev_hashjoin(d_r.data_ptr(), nbr, d_s.data_ptr(), nbs, "1", d_o.data_ptr(), nbs, ws.data_ptr(), wsb, sp)
    e1.record(); torch.cuda.synchronize()
    if it >= 2: ts.append(e0.elapsed_time(e1))
rep = dbt.stage_report(); ms = sum(ts) / len(ts)
probe = rep["hash_probe"][0]
print(f"R={nr} S={ns} kind={kind} D={D}: nres={k} sel={k/ns:.3f} total {ms:.2f} ms = {ns/ms/1e6:.2f} G probe rows/s; stages(ms) " +
      ", ".join(f"{a}={b[0]:.2f}" for a, b in rep.items()) + f"; probe kernel {(8+4*k/ns)*ns/probe/1e6/peak:.3f} of HBM peak (8+4s B/row)")
