"""Run under torchrun on >= 2 GPUs: the C++ multi-GPU layer (csrc/dist.cu through dbt.Dist: shared-memory control
block, gather+push over peer memory, pipelined key sub-ranges) against the single-node CANON oracle.  Every output ROW
is compared with all of its 140 bytes (the ranks' images concatenate with a partial block at each rank boundary, so
rows, not blocks, are the unit).  torch.distributed is only used to agree on a session name and to collect the ranks'
outputs for the comparison.  Used by tests/test_gpu_dist.py; prints DIST_CHECK_PASSED."""
import importlib
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as orc  # noqa: E402  (checker)

BLOCK = 14016


def rows_of_image(t, out_rows):
    nb = (out_rows + 99) // 100
    img = orc.as_blocks(t[: nb * BLOCK].cpu().numpy().copy())
    assert int(img["nreserved"].sum()) == out_rows, (int(img["nreserved"].sum()), out_rows)
    if nb:
        assert (img["blockid"] == np.arange(nb)).all() and (img["valid"] == 1).all()  # CANON headers, rank-local numbering
    return orc.rows_of(img).copy()


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    dbt = importlib.import_module("database-technology-algorithms_b200")
    tok = [f"{os.environ.get('MASTER_PORT', '0')}_{os.getpid()}_{int(time.time() * 1e3) % 100000}"]
    dist.broadcast_object_list(tok, src=0)
    d = dbt.Dist(tok[0], rank, world, lr)
    nb_local = int(os.environ.get("DIST_CHECK_BLOCKS", "200"))
    f1, f2 = orc.gen_ref(11, nb_local * world, num_mod=30000)
    rng = np.random.default_rng(5)
    rag = f1.copy()  # ragged inputs: partial and empty blocks anywhere (the push reads rows through the slot list)
    rag["nreserved"] = rng.integers(0, 101, size=len(rag)).astype(np.uint32)
    to_dev = lambda b: torch.from_numpy(np.ascontiguousarray(b).view(np.uint8).reshape(-1).copy()).to(dev)
    sl = slice(rank * nb_local, (rank + 1) * nb_local)
    t1, t2, t1r = to_dev(f1[sl]), to_dev(f2[sl]), to_dev(rag[sl])
    cap = nb_local * world + 8
    out = torch.empty(cap * BLOCK, dtype=torch.uint8, device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    res = {}
    for rep, q in enumerate((0, 1, 3)):  # automatic, one region, three key sub-ranges per owner (pipelined); buffers are reused
        d.set_sub_ranges(q)
        for field in ("1", "0", "2", "3"):
            for op in ("sort", "dedup"):
                for name, src in (("", t1), ("_rag", t1r)):
                    n, m = d.sort(src.data_ptr(), nb_local, field, op == "dedup", out.data_ptr(), cap, sp)
                    res[f"{op}{field}{name}_q{q}"] = rows_of_image(out, n)
            if rep == 0:
                k = d.hashjoin(t1.data_ptr(), nb_local, t2.data_ptr(), nb_local, field, out.data_ptr(), cap, sp)
                res[f"hashjoin{field}"] = rows_of_image(out, k)
                info = d.mergejoin(t1.data_ptr(), nb_local, t2.data_ptr(), nb_local, field, out.data_ptr(), cap, sp)
                res[f"mergejoin{field}"] = rows_of_image(out, info["nres"])
    st = d.stats()
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    ok = True
    if rank == 0:
        want = {}
        for field in ("1", "0", "2", "3"):
            for name, src in (("", f1), ("_rag", rag)):
                want[f"sort{field}{name}"] = (orc.rows_of(orc.sort(src, field)), False)
                want[f"dedup{field}{name}"] = (orc.rows_of(orc.dedup(src, field)), False)
            # u32 keys: replicated build keys => the ranks' outputs concatenate in S file order (exact compare);
            # str / composite keys: hash partition => compare as multisets of rows
            want[f"hashjoin{field}"] = (orc.rows_of(orc.hashjoin(f1, f2, field)), field in ("2", "3"))
            want[f"mergejoin{field}"] = (orc.rows_of(orc.mergejoin(f1, f2, field)[0]), False)
        for name in sorted(res):
            base = name.split("_q")[0]
            w, as_set = want[base]
            got = np.concatenate([g[name] for g in gathered])
            if as_set:
                same = len(got) == len(w) and sorted(x.tobytes() for x in got) == sorted(x.tobytes() for x in w)
            else:
                same = got.tobytes() == w.tobytes()  # every row, all 140 bytes, in the global order
            print(f"{name}: {'OK' if same else 'MISMATCH'} ({len(got)} rows; per rank {[len(g[name]) for g in gathered]})")
            ok = ok and same
        print("last op stats", st)
        print("DIST_CHECK_PASSED" if ok else "DIST_CHECK_FAILED")
    d.barrier()
    d.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
