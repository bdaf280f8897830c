import importlib, sys, os
sys.path.insert(0, "/root/repo")
import torch
dbt = importlib.import_module("database-technology-algorithms_b200"); L = dbt.lib()
BB=14016; nr, ns = 100_000_000, 400_000_000; nbr, nbs = nr//100, ns//100
sp = torch.cuda.current_stream().cuda_stream
d_r = torch.empty(nbr*BB, dtype=torch.uint8, device="cuda"); d_s = torch.empty(nbs*BB, dtype=torch.uint8, device="cuda"); d_o = torch.empty(nbs*BB, dtype=torch.uint8, device="cuda")
# kind 0 => keys are bij32 of small integers: spread over the full 32-bit range; R and S share the key pool
dbt.check(L.dbt_gen_syn(5, nr, nr, 0, 0, nr, 0, d_r.data_ptr(), sp)); dbt.check(L.dbt_gen_syn(5, ns, 2*nr, 0, 0, ns, 0, d_s.data_ptr(), sp))
wsb = dbt.dev_ws_bytes(dbt.OP_HASHJOIN, nbr, nbs, "1"); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
L.dbt_stage_timing_enable(1)
for env in ({}, {"DBT_JOIN_NO_SLICES": "1"}):
    for k, v in env.items(): os.environ[k] = v
    for it in range(3):
        torch.cuda.synchronize(); L.dbt_stage_timing_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
        k = dbt.dev_hashjoin(d_r.data_ptr(), nbr, d_s.data_ptr(), nbs, "1", d_o.data_ptr(), nbs, ws.data_ptr(), wsb, sp)
        e1.record(); torch.cuda.synchronize()
    print(env, "nres", k, "total", round(e0.elapsed_time(e1), 2), {a: round(b[0], 2) for a, b in dbt.stage_report().items()})
