"""BASELINE.json configs[2..4] at their stated scale on N GPUs (torchrun), device scope, synthetic inputs in HBM:
  cfg2  MergeSort field=str on  N x rows_str records           (5-letter strings + "Hola", like main.cpp)
  cfg3  HashJoin  field=num, R = 100M x S = 1B records total    (uniform and skewed S), hash-partitioned
  cfg4  MergeJoin field=num+str, 2 x 500M records total
Prints one JSON line per config from rank 0 (max-over-ranks CUDA-event time)."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
dbt = importlib.import_module("database-technology-algorithms_b200"); L = dbt.lib()
dmod = importlib.import_module("database-technology-algorithms_b200.dist")
ops = dmod.LocalOps(dev); d = dmod.DistOps(ops)
BB = 14016; sp = torch.cuda.current_stream().cuda_stream
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
only = sys.argv[2].split(",") if len(sys.argv) > 2 else ["2", "3", "4"]

def gen(seed, n_total, U, kind, rows):
    nb = rows // 100
    img = torch.empty(nb * BB, dtype=torch.uint8, device=dev)
    dbt.check(L.dbt_gen_syn(seed, n_total, U, kind, rank * rows, rows, 0, img.data_ptr(), sp))
    return img, nb

def timed(fn, reps=3):
    best = None; res = None
    for it in range(reps + 1):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); res = fn(); e1.record(); dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it: best = t.item() if best is None else min(best, t.item())
    return best, res

def report(name, ms, units, extra):
    if rank == 0:
        print(json.dumps({"config": name, "n_gpus": world, "ms": ms, "units_per_s": units / (ms * 1e-3), **extra}), flush=True)

# ---- cfg2: str sort
if "2" in only:
    rows = int(100_000_000 * scale) // 100 * 100
    img, nb = gen(11, rows * world, rows * world, 1, rows)
    ms, (out, info) = timed(lambda: d.sort(img, nb, "2"))
    tot = d.total(info["out_rows"])
    report("configs[2] MergeSort field=str", ms, rows * world, {"records": rows * world, "rows_out": tot, "ok": tot == rows * world,
           "exchange": d.last_exchange.get("mode")})
    del img, out; torch.cuda.empty_cache()

# ---- cfg3: hash join R=100M x S=1B total
nr_tot, ns_tot = int(100_000_000 * scale), int(1_000_000_000 * scale)
rr, rs = nr_tot // world // 100 * 100, ns_tot // world // 100 * 100
for kind, label in (((1, "uniform"), (2, "skewed")) if "3" in only else ()):
    r_img, nbr = gen(7, rr * world, nr_tot, 1, rr)
    s_img, nbs = gen(9, rs * world, nr_tot, kind, rs)
    ms, (out, info) = timed(lambda: d.hashjoin(r_img, nbr, s_img, nbs, "1"))
    tot = d.total(info["out_rows"])
    report(f"configs[3] HashJoin field=num {label}", ms, rs * world, {"R": rr * world, "S": rs * world, "nres": tot,
           "selectivity": tot / (rs * world), "exchange": d.last_exchange.get("mode")})
    del r_img, s_img, out; torch.cuda.empty_cache()

if "4" in only:
    # ---- cfg4: merge join composite key 2 x 500M total
    n_tot = int(500_000_000 * scale); rows = n_tot // world // 100 * 100
    r_img, nbr = gen(21, rows * world, 150_000_000, 1, rows)
    s_img, nbs = gen(21 ^ 0x5EED, rows * world, 150_000_000, 3, rows)  # ~half of S's composite keys exist in R
    ms, (out, info) = timed(lambda: d.mergejoin(r_img, nbr, s_img, nbs, "3"), reps=2)
    tot = d.total(info["out_rows"])
    report("configs[4] MergeJoin field=num+str", ms, 2 * rows * world, {"R": rows * world, "S": rows * world, "nres": tot})
dist.barrier(); dist.destroy_process_group()
