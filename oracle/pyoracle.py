"""Python face of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module (it is the checker, never the product).

Two oracles:
  * CANON -- ``oracle/dbt_oracle.c`` (liboracle.so), the defect-free CPU restatement;
  * REF   -- ``oracle/_ref/ref_runner``, the untouched reference operators compiled from
             /root/reference by ``oracle/Makefile`` (run per call in a scratch directory, in a
             child process, because the reference uses fixed file names in CWD and exit(0)).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BLOCK_BYTES = 14016
RECORD_BYTES = 140
RPB = 100

# numpy view of the on-disk layout (reference: dbtproj.h:20-38; offsets from SURVEY.md F4)
RECORD_DT = np.dtype(
    {
        # "pad" covers the 3 alignment bytes so that numpy copies move all 140 bytes
        "names": ["recid", "num", "str", "valid", "pad", "dummy1", "dummy2"],
        "formats": ["<u4", "<u4", "V120", "u1", "V3", "<u4", "<u4"],
        "offsets": [0, 4, 8, 128, 129, 132, 136],
        "itemsize": RECORD_BYTES,
    }
)
BLOCK_DT = np.dtype(
    {
        "names": ["blockid", "nreserved", "entries", "valid", "misc", "pad", "dummy"],
        "formats": ["<u4", "<u4", (RECORD_DT, (RPB,)), "u1", "u1", "V2", "<u4"],
        "offsets": [0, 4, 8, 14008, 14009, 14010, 14012],
        "itemsize": BLOCK_BYTES,
    }
)


def build(force: bool = False) -> None:
    """Compile liboracle.so (and _ref/ref_runner when /root/reference is present)."""
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "dbt_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    ref = os.path.join(HERE, "_ref", "ref_runner")
    if os.path.exists("/root/reference/DatabaseProject.cpp") and (force or not os.path.exists(ref)):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(os.path.join(HERE, "liboracle.so"))
        L = _lib
        vp, i64, u64, u32, ci = C.c_void_p, C.c_int64, C.c_uint64, C.c_uint32, C.c_int
        L.orc_count_rows.restype = i64
        L.orc_count_rows.argtypes = [vp, i64]
        L.orc_sort.restype = i64
        L.orc_sort.argtypes = [vp, i64, ci, vp]
        L.orc_dedup.restype = i64
        L.orc_dedup.argtypes = [vp, i64, ci, vp]
        L.orc_mergejoin.restype = None
        L.orc_mergejoin.argtypes = [vp, i64, vp, i64, ci, vp, vp, vp, C.POINTER(i64)]
        L.orc_hashjoin.restype = i64
        L.orc_hashjoin.argtypes = [vp, i64, vp, i64, ci, vp]
        L.orc_innerjoin_pairs.restype = i64
        L.orc_innerjoin_pairs.argtypes = [vp, i64, vp, i64, ci, vp]
        L.orc_sort_counters.restype = None
        L.orc_sort_counters.argtypes = [i64, i64, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
        L.orc_dedup_nios.restype = i64
        L.orc_dedup_nios.argtypes = [i64, i64, i64]
        L.orc_hashjoin_nios.restype = i64
        L.orc_hashjoin_nios.argtypes = [i64, i64, i64, i64]
        L.orc_gen_ref.restype = None
        L.orc_gen_ref.argtypes = [C.c_uint, i64, ci, u32, vp, vp]
        L.orc_gen_syn.restype = None
        L.orc_gen_syn.argtypes = [u64, u64, u64, ci, u64, u64, u32, vp]
        L.orc_syn_num.restype = u32
        L.orc_syn_num.argtypes = [u64, u64, u64, ci, u64]
        L.orc_is_sorted.restype = ci
        L.orc_is_sorted.argtypes = [vp, i64, ci]
        L.orc_canonicalise_ties.restype = None
        L.orc_canonicalise_ties.argtypes = [vp, i64, ci]
    return _lib


def _p(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def _fld(field) -> int:
    return ord(field) if isinstance(field, str) else int(field)


def new_blocks(nblocks: int) -> np.ndarray:
    return np.zeros(max(int(nblocks), 0), dtype=BLOCK_DT)


def as_blocks(buf) -> np.ndarray:
    """View raw bytes (bytes / uint8 array) as a block array."""
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf.view(np.uint8).reshape(-1)
    assert a.size % BLOCK_BYTES == 0, "not a whole number of blocks"
    return a.view(BLOCK_DT)


def rows_of(blocks: np.ndarray, nrows: int | None = None) -> np.ndarray:
    """Live rows of an image in file order (by nreserved, or the first ``nrows`` slots when the
    headers cannot be trusted, as for REF join outputs)."""
    if nrows is not None:
        return blocks["entries"].reshape(-1)[:nrows]
    nres = np.minimum(blocks["nreserved"], RPB)
    mask = np.arange(RPB)[None, :] < nres[:, None]
    return blocks["entries"][mask]


def count_rows(blocks: np.ndarray) -> int:
    return int(lib().orc_count_rows(_p(blocks), len(blocks)))


def gen_ref(seed: int, nblocks: int, tail_mode: int = 0, num_mod: int = 0, two: bool = True):
    """G_ref: reference main.cpp:41-77 with srand(seed)."""
    f1 = new_blocks(nblocks)
    f2 = new_blocks(nblocks) if two else None
    lib().orc_gen_ref(seed, nblocks, tail_mode, num_mod, _p(f1), _p(f2) if two else None)
    return (f1, f2) if two else f1


def gen_syn(seed: int, n_total: int, U: int, kind: int, row0: int = 0, nrows: int | None = None, recid0: int = 0):
    nrows = n_total - row0 if nrows is None else nrows
    out = new_blocks((nrows + RPB - 1) // RPB)
    lib().orc_gen_syn(seed, n_total, U, kind, row0, nrows, recid0, _p(out))
    return out


def sort(blocks: np.ndarray, field) -> np.ndarray:
    n = count_rows(blocks)
    out = new_blocks((n + RPB - 1) // RPB)
    lib().orc_sort(_p(blocks), len(blocks), _fld(field), _p(out))
    return out


def dedup(blocks: np.ndarray, field) -> np.ndarray:
    n = count_rows(blocks)
    out = new_blocks((n + RPB - 1) // RPB)
    u = lib().orc_dedup(_p(blocks), len(blocks), _fld(field), _p(out))
    return out[: (u + RPB - 1) // RPB]


def mergejoin(r: np.ndarray, s: np.ndarray, field):
    nr, ns = count_rows(r), count_rows(s)
    ur = new_blocks((nr + RPB - 1) // RPB)
    us = new_blocks((ns + RPB - 1) // RPB)
    out = new_blocks((min(nr, ns) + RPB - 1) // RPB)
    res = (C.c_int64 * 4)()
    lib().orc_mergejoin(_p(r), len(r), _p(s), len(s), _fld(field), _p(ur), _p(us), _p(out), res)
    k, a, b, reads = [int(x) for x in res]
    nb = lambda n: (n + RPB - 1) // RPB
    return out[: nb(k)], ur[: nb(a)], us[: nb(b)], {"nres": k, "nunique_R": a, "nunique_S": b, "later_reads": reads}


def hashjoin(r: np.ndarray, s: np.ndarray, field) -> np.ndarray:
    k = lib().orc_hashjoin(_p(r), len(r), _p(s), len(s), _fld(field), None)
    out = new_blocks((k + RPB - 1) // RPB)
    lib().orc_hashjoin(_p(r), len(r), _p(s), len(s), _fld(field), _p(out))
    return out


def innerjoin_pairs(r: np.ndarray, s: np.ndarray, field) -> np.ndarray:
    """(recid_R, recid_S) for every matching row pair (extension)."""
    k = lib().orc_innerjoin_pairs(_p(r), len(r), _p(s), len(s), _fld(field), None)
    out = np.zeros((k, 2), dtype=np.uint32)
    lib().orc_innerjoin_pairs(_p(r), len(r), _p(s), len(s), _fld(field), _p(out))
    return out


def sort_counters(B: int, M: int):
    a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
    lib().orc_sort_counters(B, M, C.byref(a), C.byref(b), C.byref(c))
    return {"nsorted_segs": a.value, "npasses": b.value, "nios": c.value}


def dedup_nios(B, M, nunique):
    return int(lib().orc_dedup_nios(B, M, nunique))


def hashjoin_nios(BR, BS, M, nres):
    return int(lib().orc_hashjoin_nios(BR, BS, M, nres))


def mergejoin_nios(BR, BS, M, info):
    """CANON nios of MergeJoin (Appendix B): both dedups + 2 first reads + later reads + writes."""
    return (
        dedup_nios(BR, M, info["nunique_R"])
        + dedup_nios(BS, M, info["nunique_S"])
        + 2
        + info["later_reads"]
        + (info["nres"] + RPB - 1) // RPB
    )


def is_sorted(blocks, field) -> bool:
    return bool(lib().orc_is_sorted(_p(blocks), len(blocks), _fld(field)))


def canonicalise_ties(blocks, field) -> np.ndarray:
    out = blocks.copy()
    lib().orc_canonicalise_ties(_p(out), len(out), _fld(field))
    return out


# ---------------------------------------------------------------------------------------------
# REF: the untouched reference, through oracle/_ref/ref_runner
# ---------------------------------------------------------------------------------------------
REF_RUNNER = os.path.join(HERE, "_ref", "ref_runner")


def ref_available() -> bool:
    return os.path.exists(REF_RUNNER) and os.access(REF_RUNNER, os.X_OK)


def scratch_root() -> str:
    return "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()


def run_ref(op: str, field: str, nmem: int, r: np.ndarray, s: np.ndarray | None = None, keep_dir: bool = False,
            timeout: float = 3600.0):
    """Run one REF operator on in-memory images.  Returns (info dict, out blocks, extra files).

    ``info['seconds']`` is the wall time of the operator call alone (steady_clock inside the child).
    """
    assert ref_available(), "oracle/_ref/ref_runner is not built (run `make -C oracle`)"
    d = tempfile.mkdtemp(prefix="dbtref_", dir=scratch_root())
    try:
        r.tofile(os.path.join(d, "file.bin"))
        args = [REF_RUNNER, op, field, str(nmem), "file.bin"]
        if op in ("mjoin", "hjoin"):
            s.tofile(os.path.join(d, "file2.bin"))
            args.append("file2.bin")
        if op != "sort":
            args.append("out.bin")
        p = subprocess.run(args, cwd=d, capture_output=True, text=True, timeout=timeout)
        line = [l for l in p.stdout.splitlines() if l.startswith("{")]
        if not line:
            return {"exit": p.returncode, "stdout": p.stdout, "stderr": p.stderr}, None, {}
        info = json.loads(line[-1])
        info["exit"] = p.returncode
        info["stdout"] = p.stdout
        out = as_blocks(np.fromfile(os.path.join(d, info["outfile"]), dtype=np.uint8))
        extra = {}
        if op == "mjoin":
            for name in ("1outfile.bin", "2outfile.bin"):
                extra[name] = as_blocks(np.fromfile(os.path.join(d, name), dtype=np.uint8))
        return info, out, extra
    finally:
        if not keep_dir:
            shutil.rmtree(d, ignore_errors=True)
