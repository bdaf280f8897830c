"""BASELINE configs[0]: MergeSort field=num on a 1M-record dbtproj file, nmem_blocks=64 -- the file-based drop-in
entry point (libdbt_b200.so, called through its mangled name like main.o would) against the untouched reference
(oracle/_ref/ref_runner) on the same box, files in tmpfs.  Also EliminateDuplicates / HashJoin / MergeJoin."""
import ctypes as C, importlib, os, sys, tempfile, time, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as orc  # baseline leg
dbt = importlib.import_module("database-technology-algorithms_b200"); L = dbt.lib()
nblocks = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
f1, f2 = orc.gen_ref(42, nblocks)
d = tempfile.mkdtemp(prefix="dbtcfg0_", dir=orc.scratch_root()); os.chdir(d)
f1.tofile("file.bin"); f2.tofile("file2.bin")
def entry(name):
    f = getattr(L, dbt.CXX_ENTRY_POINTS[name]); f.restype = None; return f
a, b, c = C.c_uint(), C.c_uint(), C.c_uint(); out = C.create_string_buffer(64)
def t(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
    return best
res = {}
res["MergeSort"] = t(lambda: entry("MergeSort")(b"file.bin", C.c_ubyte(ord("1")), None, 64, out, C.byref(a), C.byref(b), C.byref(c)))
ours = (a.value, b.value, c.value, out.value.decode())
res["EliminateDuplicates"] = t(lambda: entry("EliminateDuplicates")(b"file.bin", C.c_ubyte(ord("1")), None, 64, b"nodup.bin", C.byref(a), C.byref(c)))
res["MergeJoin"] = t(lambda: entry("MergeJoin")(b"file.bin", b"file2.bin", C.c_ubyte(ord("1")), None, 64, b"mj.bin", C.byref(a), C.byref(c)))
res["HashJoin"] = t(lambda: entry("HashJoin")(b"file.bin", b"file2.bin", C.c_ubyte(ord("1")), None, 64, b"hj.bin", C.byref(a), C.byref(c)))
ref = {}
if orc.ref_available():
    for op, name in (("sort", "MergeSort"), ("dedup", "EliminateDuplicates"), ("mjoin", "MergeJoin"), ("hjoin", "HashJoin")):
        info, _, _ = orc.run_ref(op, "1", 64, f1, f2 if op in ("mjoin", "hjoin") else None)
        ref[name] = info["seconds"]
        if op == "sort": refc = (info["a"], info["b"], info["nios"], info["outfile"])
print(f"rows per file: {nblocks*100}; ours counters {ours}; reference counters {refc if ref else None}")
for k in res:
    print(f"{k:20s} libdbt_b200 (file->file, incl. pinned alloc, H2D, D2H, write) {res[k]*1e3:8.1f} ms   reference {ref.get(k, float('nan'))*1e3:8.1f} ms   speed-up {ref.get(k, float('nan'))/res[k]:6.1f}x")
os.chdir("/"); shutil.rmtree(d, ignore_errors=True)
