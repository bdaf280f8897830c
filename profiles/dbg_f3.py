#!/usr/bin/env python
"""Diagnostic: HashJoin field '3' with heavy multiplicities at growing sizes; product count vs numpy truth vs pair count."""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
dbt = importlib.import_module("database-technology-algorithms_b200")
from oracle import pyoracle as orc
import helpers as H

orc.build()
L = dbt.lib()
for nb, hot in ((100, 1000), (1000, 10000), (10000, 100000), (10000, 1000000)):
    r = orc.gen_syn(5, nb * 100, 1000, 1)
    s = orc.gen_syn(6, nb * 100, 1000, 1)
    for img in (r, s):
        img["entries"]["str"] = np.zeros(120, np.uint8).view("V120")[0]
    e = r["entries"].reshape(-1).copy(); e["num"][hot:] += 5000; r["entries"][:] = e.reshape(r["entries"].shape)
    e = s["entries"].reshape(-1).copy(); e["num"][hot:] += 9000; s["entries"][:] = e.reshape(s["entries"].shape)
    rn = orc.rows_of(r)["num"]; sn = orc.rows_of(s)["num"]
    per_key = np.bincount(rn, minlength=20000)
    want = int(per_key[sn].sum())
    d_r, d_s = H.to_dev(r), H.to_dev(s)
    cap = (want + 99) // 100 + 10
    d_out = H.dev_alloc(cap * H.BLOCK_BYTES)
    wsb = L.dbt_dev_hashjoin_ws_bytes(nb, nb, ord("3"), 8, cap)
    ws = H.dev_alloc(wsb)
    n = C.c_uint64()
    rc = L.dbt_dev_hashjoin(d_r.data_ptr(), nb, d_s.data_ptr(), nb, ord("3"), d_out.data_ptr(), cap, ws.data_ptr(), wsb, H.stream(), C.byref(n))
    npairs = C.c_uint64()
    wsb2 = dbt.dev_ws_bytes(dbt.OP_MERGEJOIN, nb, nb, "3")
    ws2 = H.dev_alloc(wsb2)
    rc2 = L.dbt_dev_innerjoin_pairs(d_r.data_ptr(), nb, d_s.data_ptr(), nb, ord("3"), None, 0, ws2.data_ptr(), wsb2, H.stream(), C.byref(npairs))
    print(f"nb={nb} hot={hot}: want={want} hashjoin rc={rc} n={n.value} pairs rc={rc2} npairs={npairs.value}", flush=True)
    if rc == 0 and n.value != want:
        out = H.to_host(d_out, (n.value + 99) // 100, orc)
        ids = orc.rows_of(out)["recid"]
        got_per_s = np.bincount(ids, minlength=nb * 100)
        exp_per_s = per_key[sn]
        bad = np.flatnonzero(got_per_s != exp_per_s)
        print("  rows with wrong multiplicity:", len(bad), "first:", bad[:10], "got", got_per_s[bad[:10]], "want", exp_per_s[bad[:10]])
        keys_bad = np.unique(sn[bad])
        print("  distinct keys affected:", len(keys_bad), "of", len(np.unique(sn[exp_per_s > 0])), "sample", keys_bad[:10])
        # is the shortfall per key constant across its S rows?
        k0 = keys_bad[0]
        rows_k0 = np.flatnonzero(sn == k0)
        print("  key", k0, "R multiplicity", per_key[k0], "emitted per S row:", np.unique(got_per_s[rows_k0]))
