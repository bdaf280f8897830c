"""Run under torchrun on >= 2 GPUs: the CUDA multi-GPU path (DistOps + LocalOps, fused gather->peer
exchange or NCCL all-to-all) against the single-node CANON oracle.  Used by tests/test_gpu_dist.py."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as orc  # noqa: E402  (checker)

BLOCK = 14016


def rows_of_image(t, out_rows):
    nb = (out_rows + 99) // 100
    return orc.rows_of(orc.as_blocks(t[: nb * BLOCK].cpu().numpy().copy()))["recid"].copy()


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    dmod = importlib.import_module("database-technology-algorithms_b200.dist")
    ops = dmod.LocalOps(dev)
    d = dmod.DistOps(ops, samples_per_rank=1024)
    nb_local = 200
    f1, f2 = orc.gen_ref(11, nb_local * world, num_mod=30000)
    to_dev = lambda b: torch.from_numpy(b.view(np.uint8).reshape(-1).copy()).to(dev)
    t1 = to_dev(f1[rank * nb_local:(rank + 1) * nb_local])
    t2 = to_dev(f2[rank * nb_local:(rank + 1) * nb_local])
    res = {}
    for rep in range(2):  # twice: buffer reuse across steps must be safe
        for field in ("1", "0", "2", "3"):
            for op in ("sort", "dedup"):
                out, info = getattr(d, op)(t1, nb_local, field)
                res[f"{op}{field}"] = rows_of_image(out, info["out_rows"])
            out, info = d.hashjoin(t1, nb_local, t2, nb_local, field)
            res[f"hashjoin{field}"] = rows_of_image(out, info["out_rows"])
            out, info = d.mergejoin(t1, nb_local, t2, nb_local, field)
            res[f"mergejoin{field}"] = rows_of_image(out, info["out_rows"])
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    ok = True
    if rank == 0:
        for field in ("1", "0", "2", "3"):
            checks = {
                f"sort{field}": (orc.rows_of(orc.sort(f1, field))["recid"], False),
                f"dedup{field}": (orc.rows_of(orc.dedup(f1, field))["recid"], False),
                # u32 keys: replicated build keys => the ranks' outputs concatenate in S file order (exact compare);
                # str / composite keys: hash partition => compare as sorted multisets
                f"hashjoin{field}": (orc.rows_of(orc.hashjoin(f1, f2, field))["recid"], field in ("2", "3")),
                f"mergejoin{field}": (orc.rows_of(orc.mergejoin(f1, f2, field)[0])["recid"], False),
            }
            for name, (want, as_set) in checks.items():
                got = np.concatenate([g[name] for g in gathered])
                same = np.array_equal(np.sort(got), np.sort(want)) if as_set else np.array_equal(got, want)
                print(f"{name}: {'OK' if same else 'MISMATCH'} ({len(got)} rows, exchange={d.last_exchange.get('mode')})")
                ok = ok and same
        print("DIST_CHECK_PASSED" if ok else "DIST_CHECK_FAILED")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
