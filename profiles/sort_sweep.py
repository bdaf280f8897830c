"""Times dbt_sort_pairs_u32 (device scope) for a given n; prints ms and fraction of the HBM roofline.
Usage: python profiles/sort_sweep.py <n> [iters]   (env: DBT_ONESWEEP_IMPL / _RANK / _CFG / _CTAS)"""
import ctypes as C, importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dbt = importlib.import_module("database-technology-algorithms_b200"); L = dbt.lib()
n = int(sys.argv[1]); iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
g = torch.Generator(device="cuda").manual_seed(1)
keys = torch.randint(-2**31, 2**31, (n,), dtype=torch.int32, device="cuda", generator=g)
k1, k2, v1, v2 = (torch.empty_like(keys) for _ in range(4))
wsb = L.dbt_sort_pairs_ws_bytes(n); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
alt = C.c_int(); sp = torch.cuda.current_stream().cuda_stream
L.dbt_stage_timing_enable(1)
ts = []
iota = torch.arange(n, dtype=torch.int32, device="cuda")
for it in range(iters + 2):
    k1.copy_(keys); v1.copy_(iota)
    torch.cuda.synchronize(); L.dbt_stage_timing_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dbt.check(L.dbt_sort_pairs_u32(k1.data_ptr(), k2.data_ptr(), v1.data_ptr(), v2.data_ptr(), n, 0, 32, ws.data_ptr(), wsb, sp, C.byref(alt)))
    e1.record(); torch.cuda.synchronize()
    if it >= 2: ts.append(e0.elapsed_time(e1))
rep = dbt.stage_report()
ko = k2 if alt.value else k1
c = ko[: min(n, 50_000_000)].to(torch.int64) & 0xFFFFFFFF
ok = bool((c[1:] >= c[:-1]).all().item())
if n <= 200_000_000:  # exact: the row column must be the stable argsort of the keys
    vo = v2 if alt.value else v1
    want = torch.sort(keys.to(torch.int64) & 0xFFFFFFFF, stable=True).indices
    ok = ok and bool((vo.to(torch.int64) == want).all().item())
    del want
ms = sum(ts) / len(ts); one = rep["onesweep_pass"][0] / 4
print(f"n={n} total {ms:.3f} ms  onesweep/pass {one:.3f} ms = {16*n/one/1e6/peak:.3f} of HBM peak  hist {rep['histogram'][0]:.3f} ms  sorted={ok}  env={ {k:v for k,v in os.environ.items() if k.startswith('DBT_')} }")
