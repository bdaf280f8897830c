"""Run with DBT_DEVICES=<n> (n >= 2 GPUs, one process): the four dbtproj.h entry points on several GPUs (csrc/host_multi.cu:
one host thread per GPU, dbt_dist_init_local group, outputs re-blocked at the rank boundaries) must write files that are
byte-identical to the oracle's images, with the reference's counters.  Used by tests/test_gpu_dist.py; prints
MULTI_FILE_CHECK_PASSED."""
import ctypes as C
import importlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle as orc  # noqa: E402  (checker)

dbt = importlib.import_module("database-technology-algorithms_b200")
L = dbt.lib()


def entry(name):
    f = getattr(L, dbt.CXX_ENTRY_POINTS[name])
    f.restype = None
    return f


def blocks(path):
    return orc.as_blocks(np.fromfile(path, dtype=np.uint8))


def same(a, b):
    return a.shape == b.shape and a.tobytes() == b.tobytes()


def main():
    assert int(os.environ.get("DBT_DEVICES", "0")) >= 2
    orc.build()
    d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    os.chdir(d)
    nb = int(os.environ.get("MULTI_CHECK_BLOCKS", "1500"))
    f1, f2 = orc.gen_ref(23, nb, num_mod=nb * 12)
    rng = np.random.default_rng(1)
    f1["nreserved"][rng.integers(0, nb, size=40)] = rng.integers(0, 100, size=40).astype(np.uint32)  # some partial blocks
    f1.tofile("file.bin")
    f2.tofile("file2.bin")
    ok = True

    def report(name, cond):
        nonlocal ok
        print(f"{name}: {'OK' if cond else 'MISMATCH'}", flush=True)
        ok = ok and bool(cond)

    for field in "0123":
        out = C.create_string_buffer(64)
        a, b, c = C.c_uint(), C.c_uint(), C.c_uint()
        entry("MergeSort")(b"file.bin", C.c_ubyte(ord(field)), None, C.c_uint(64), out, C.byref(a), C.byref(b), C.byref(c))
        want = orc.sort(f1, field)
        cnt = orc.sort_counters(nb, 64)
        report(f"MergeSort field {field}", same(blocks(out.value.decode()), want) and (a.value, b.value, c.value) == (cnt["nsorted_segs"], cnt["npasses"], cnt["nios"]))
    for field in "13":
        u, io = C.c_uint(), C.c_uint()
        entry("EliminateDuplicates")(b"file.bin", C.c_ubyte(ord(field)), None, C.c_uint(64), b"nodup.bin", C.byref(u), C.byref(io))
        want = orc.dedup(f1, field)
        report(f"EliminateDuplicates field {field}", same(blocks("nodup.bin"), want) and u.value == orc.count_rows(want) and io.value == orc.dedup_nios(nb, 64, u.value))
    for field in "012":  # '2' stays on one GPU (hash-partitioned output would not be in S file order)
        n, io = C.c_uint(), C.c_uint()
        entry("HashJoin")(b"file.bin", b"file2.bin", C.c_ubyte(ord(field)), None, C.c_uint(64), b"oh.bin", C.byref(n), C.byref(io))
        want = orc.hashjoin(f1, f2, field)
        report(f"HashJoin field {field}", same(blocks("oh.bin"), want) and n.value == orc.count_rows(want) and io.value == orc.hashjoin_nios(nb, nb, 64, n.value))
    for field in "13":
        n, io = C.c_uint(), C.c_uint()
        entry("MergeJoin")(b"file.bin", b"file2.bin", C.c_ubyte(ord(field)), None, C.c_uint(100), b"om.bin", C.byref(n), C.byref(io))
        want, ur, us, info = orc.mergejoin(f1, f2, field)
        report(f"MergeJoin field {field}", same(blocks("om.bin"), want) and same(blocks("1outfile.bin"), ur) and same(blocks("2outfile.bin"), us)
               and n.value == info["nres"] and io.value == orc.mergejoin_nios(nb, nb, 100, info))
    print("MULTI_FILE_CHECK_PASSED" if ok else "MULTI_FILE_CHECK_FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
