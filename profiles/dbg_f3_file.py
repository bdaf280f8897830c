#!/usr/bin/env python
"""Diagnostic: HashJoin() file entry point, field '3', output ~10x S (the capacity retry path)."""
import ctypes as C, importlib, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
dbt = importlib.import_module("database-technology-algorithms_b200")
from oracle import pyoracle as orc
orc.build()
L = dbt.lib()
d = tempfile.mkdtemp(dir="/dev/shm"); os.chdir(d)
nb = 10000
r = orc.gen_syn(5, nb * 100, 1000, 1); s = orc.gen_syn(6, nb * 100, 1000, 1)
for img in (r, s):
    img["entries"]["str"] = np.zeros(120, np.uint8).view("V120")[0]
e = r["entries"].reshape(-1).copy(); e["num"][100000:] += 5000; r["entries"][:] = e.reshape(r["entries"].shape)
e = s["entries"].reshape(-1).copy(); e["num"][100000:] += 9000; s["entries"][:] = e.reshape(s["entries"].shape)
r.tofile("r.bin"); s.tofile("s.bin")
rn = orc.rows_of(r)["num"]; sn = orc.rows_of(s)["num"]
per_key = np.bincount(rn, minlength=20000); want = int(per_key[sn].sum())
f = getattr(L, dbt.CXX_ENTRY_POINTS["HashJoin"]); f.restype = None
for rep in range(3):
    a, c = C.c_uint(), C.c_uint()
    f(b"r.bin", b"s.bin", C.c_ubyte(ord("3")), None, C.c_uint(64), b"out.bin", C.byref(a), C.byref(c))
    out = orc.as_blocks(np.fromfile("out.bin", dtype=np.uint8))
    print(f"rep {rep}: want {want} nres {a.value} rows in file {orc.count_rows(out)} last_error={L.dbt_last_error()}", flush=True)
    if a.value != want:
        ids = orc.rows_of(out)["recid"]
        got_per_s = np.bincount(ids, minlength=nb * 100); exp = per_key[sn]
        bad = np.flatnonzero(got_per_s != exp)
        print("  bad S rows:", len(bad), bad[:8], got_per_s[bad[:8]], exp[bad[:8]], "max bad row", bad.max(), "min", bad.min())
