"""database-technology-algorithms_b200 -- B200-native tuple operators behind the dbtproj.h API.

The product is ``libdbt_b200.so`` (hand-written sm_100a CUDA + a C++ host layer, built in-tree by
``build.py``).  This module is only the ctypes face of its ``extern "C"`` boundary
(``include/dbt_b200.h``) for the tests and ``bench.py``; PyTorch, where used by callers, is
plumbing for device memory and streams.  There is no CPU fallback: if the shared library is
missing, or no CUDA device is visible, every operator raises.

Import with ``importlib.import_module("database-technology-algorithms_b200")`` (the directory
name carries a hyphen).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DBT_LIB") or os.path.join(HERE, "libdbt_b200.so")  # DBT_LIB: an A/B build of the same library

BLOCK_BYTES = 14016
RECORD_BYTES = 140
RPB = 100
OP_SORT, OP_DEDUP, OP_MERGEJOIN, OP_HASHJOIN = 0, 1, 2, 3

# every symbol include/dbt_b200.h declares (checked by tests/test_abi.py)
C_ABI_SYMBOLS = [
    "dbt_last_error", "dbt_abi_version", "dbt_device_count",
    "dbt_sort_counters", "dbt_dedup_nios", "dbt_hashjoin_nios", "dbt_mergejoin_nios",
    "dbt_sort_pairs_ws_bytes", "dbt_sort_pairs_u32", "dbt_gather_records",
    "dbt_dev_extract_keys_u32", "dbt_dev_partition_rows", "dbt_dev_partition_ws_bytes",
    "dbt_dev_ws_bytes", "dbt_dev_ws_bytes_kw", "dbt_dev_hashjoin_ws_bytes", "dbt_dev_mergesort", "dbt_dev_dedup", "dbt_dev_mergejoin", "dbt_dev_hashjoin", "dbt_dev_innerjoin_pairs", "dbt_dev_semijoin_keys",
    "dbt_host_mergesort", "dbt_host_dedup", "dbt_host_mergejoin", "dbt_host_hashjoin",
    "dbt_host_mergesort_begin", "dbt_host_dedup_begin", "dbt_host_mergejoin_begin", "dbt_host_hashjoin_begin",
    "dbt_host_job_wait", "dbt_host_job_slots", "dbt_host_trim", "dbt_host_set_chunk_blocks", "dbt_host_ooc_stats",
    "dbt_host_alloc", "dbt_host_free", "dbt_gen_syn",
    "dbt_dist_init", "dbt_dist_init_local", "dbt_dist_destroy", "dbt_dist_rank", "dbt_dist_world", "dbt_dist_barrier", "dbt_dist_trim",
    "dbt_dist_set_sub_ranges", "dbt_dist_allgather_host", "dbt_dist_sort", "dbt_dist_hashjoin", "dbt_dist_mergejoin",
    "dbt_dist_stats", "dbt_dist_selftest_host",
    "dbt_stage_timing_enable", "dbt_stage_timing_reset", "dbt_stage_count", "dbt_stage_name", "dbt_stage_ms",
    "dbt_stage_launches", "dbt_kernel_launches",
]
# the four drop-in entry points, C++ linkage (Itanium mangling), as main.cpp imports them
CXX_ENTRY_POINTS = {
    "MergeSort": "_Z9MergeSortPchP7block_tjS_PjS2_S2_",
    "EliminateDuplicates": "_Z19EliminateDuplicatesPchP7block_tjS_PjS2_",
    "MergeJoin": "_Z9MergeJoinPcS_hP7block_tjS_PjS2_",
    "HashJoin": "_Z8HashJoinPcS_hP7block_tjS_PjS2_",
}


class DbtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"dbt error {code}: {msg}")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    """Load libdbt_b200.so (raises if it has not been built: there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python database-technology-algorithms_b200/build.py` "
            "(there is no CPU fallback)"
        )
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, ci, sz = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_size_t
    pu64 = C.POINTER(C.c_uint64)
    L.dbt_last_error.restype = C.c_char_p
    L.dbt_sort_counters.argtypes = [u64, u32, pu64, pu64, pu64]
    L.dbt_dedup_nios.restype = u64
    L.dbt_dedup_nios.argtypes = [u64, u32, u64]
    L.dbt_hashjoin_nios.restype = u64
    L.dbt_hashjoin_nios.argtypes = [u64, u64, u32, u64]
    L.dbt_mergejoin_nios.restype = u64
    L.dbt_mergejoin_nios.argtypes = [u64, u64, u32, pu64]
    L.dbt_sort_pairs_ws_bytes.restype = sz
    L.dbt_sort_pairs_ws_bytes.argtypes = [u64]
    L.dbt_sort_pairs_u32.argtypes = [vp, vp, vp, vp, u64, ci, ci, vp, sz, vp, C.POINTER(ci)]
    L.dbt_gather_records.argtypes = [vp, vp, vp, u64, vp, vp]
    L.dbt_dev_ws_bytes.restype = sz
    L.dbt_dev_ws_bytes.argtypes = [ci, u64, u64, ci]
    L.dbt_dev_ws_bytes_kw.restype = sz
    L.dbt_dev_ws_bytes_kw.argtypes = [ci, u64, u64, ci, u32]
    L.dbt_dev_hashjoin_ws_bytes.restype = sz
    L.dbt_dev_hashjoin_ws_bytes.argtypes = [u64, u64, ci, u32, u64]
    L.dbt_dev_mergesort.argtypes = [vp, u64, ci, vp, vp, sz, vp, pu64]
    L.dbt_dev_dedup.argtypes = [vp, u64, ci, vp, vp, sz, vp, pu64, pu64]
    L.dbt_dev_mergejoin.argtypes = [vp, u64, vp, u64, ci, vp, vp, vp, vp, sz, vp, pu64]
    L.dbt_dev_hashjoin.argtypes = [vp, u64, vp, u64, ci, vp, u64, vp, sz, vp, pu64]
    L.dbt_dev_semijoin_keys.argtypes = [vp, u64, vp, u64, ci, vp, u64, vp, sz, vp, pu64]
    L.dbt_dev_innerjoin_pairs.argtypes = [vp, u64, vp, u64, ci, vp, u64, vp, sz, vp, pu64]
    L.dbt_dev_extract_keys_u32.argtypes = [vp, u64, ci, vp, vp, sz, vp, pu64]
    L.dbt_dev_partition_rows.argtypes = [vp, u64, ci, C.POINTER(u32), u32, vp, pu64, vp, sz, vp]
    L.dbt_dev_partition_ws_bytes.restype = sz
    L.dbt_dev_partition_ws_bytes.argtypes = [u64]
    L.dbt_host_mergesort.argtypes = [vp, u64, ci, vp, ci, pu64]
    L.dbt_host_dedup.argtypes = [vp, u64, ci, vp, ci, pu64, pu64]
    L.dbt_host_mergejoin.argtypes = [vp, u64, vp, u64, ci, vp, vp, vp, ci, pu64]
    L.dbt_host_hashjoin.argtypes = [vp, u64, vp, u64, ci, vp, u64, ci, pu64]
    L.dbt_host_mergesort_begin.argtypes = [ci, vp, u64, ci, vp, ci]
    L.dbt_host_dedup_begin.argtypes = [ci, vp, u64, ci, vp, ci]
    L.dbt_host_mergejoin_begin.argtypes = [ci, vp, u64, vp, u64, ci, vp, vp, vp, ci]
    L.dbt_host_hashjoin_begin.argtypes = [ci, vp, u64, vp, u64, ci, vp, u64, ci]
    L.dbt_host_job_wait.argtypes = [ci, pu64]
    L.dbt_host_job_slots.argtypes = []
    L.dbt_host_trim.argtypes = []
    L.dbt_host_set_chunk_blocks.argtypes = [u64]
    L.dbt_host_ooc_stats.argtypes = [pu64]
    L.dbt_host_alloc.argtypes = [C.POINTER(vp), sz]
    L.dbt_host_free.argtypes = [vp]
    L.dbt_gen_syn.argtypes = [u64, u64, u64, ci, u64, u64, u32, vp, vp]
    L.dbt_dist_init.argtypes = [C.c_char_p, ci, ci, ci, C.POINTER(vp)]
    L.dbt_dist_init_local.argtypes = [ci, C.POINTER(ci), C.POINTER(vp)]
    L.dbt_dist_destroy.argtypes = [vp]
    L.dbt_dist_rank.argtypes = [vp]
    L.dbt_dist_world.argtypes = [vp]
    L.dbt_dist_barrier.argtypes = [vp]
    L.dbt_dist_trim.argtypes = [vp]
    L.dbt_dist_set_sub_ranges.argtypes = [vp, u32]
    L.dbt_dist_allgather_host.argtypes = [vp, vp, sz, vp]
    L.dbt_dist_sort.argtypes = [vp, vp, u64, ci, ci, vp, u64, vp, pu64, pu64]
    L.dbt_dist_hashjoin.argtypes = [vp, vp, u64, vp, u64, ci, vp, u64, vp, pu64]
    L.dbt_dist_mergejoin.argtypes = [vp, vp, u64, vp, u64, ci, vp, u64, vp, pu64]
    L.dbt_dist_stats.argtypes = [vp, C.POINTER(C.c_double)]
    L.dbt_dist_selftest_host.argtypes = [C.c_char_p, ci, ci, pu64]
    L.dbt_stage_timing_enable.argtypes = [ci]
    L.dbt_stage_timing_enable.restype = None
    L.dbt_stage_timing_reset.restype = None
    L.dbt_stage_name.restype = C.c_char_p
    L.dbt_stage_name.argtypes = [ci]
    L.dbt_stage_ms.restype = C.c_double
    L.dbt_stage_ms.argtypes = [ci]
    L.dbt_stage_launches.restype = u64
    L.dbt_stage_launches.argtypes = [ci]
    L.dbt_kernel_launches.restype = u64
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise DbtError(rc, lib().dbt_last_error().decode(errors="replace"))


def _fld(field) -> int:
    return ord(field) if isinstance(field, str) else int(field)


def sort_counters(nblocks: int, nmem_blocks: int) -> dict:
    a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    check(lib().dbt_sort_counters(nblocks, nmem_blocks, C.byref(a), C.byref(b), C.byref(c)))
    return {"nsorted_segs": a.value, "npasses": b.value, "nios": c.value}


def dev_ws_bytes(op: int, nblocks_r: int, nblocks_s: int, field, kw: int = 8) -> int:
    return int(lib().dbt_dev_ws_bytes_kw(op, nblocks_r, nblocks_s, _fld(field), kw))


def stage_report() -> dict:
    """{stage name: (accumulated ms, kernel launches)} since the last reset (timing must be enabled)."""
    L = lib()
    out = {}
    for i in range(L.dbt_stage_count()):
        ms, n = L.dbt_stage_ms(i), L.dbt_stage_launches(i)
        if ms or n:
            out[L.dbt_stage_name(i).decode()] = (ms, int(n))
    return out


# ---- thin device-scope wrappers over raw pointers (ints) ------------------------------------
def dev_mergesort(d_in: int, nblocks: int, field, d_out: int, d_ws: int, ws_bytes: int, stream: int = 0) -> int:
    n = C.c_uint64()
    check(lib().dbt_dev_mergesort(d_in, nblocks, _fld(field), d_out, d_ws, ws_bytes, stream, C.byref(n)))
    return n.value


def dev_dedup(d_in: int, nblocks: int, field, d_out: int, d_ws: int, ws_bytes: int, stream: int = 0):
    n, u = C.c_uint64(), C.c_uint64()
    check(lib().dbt_dev_dedup(d_in, nblocks, _fld(field), d_out, d_ws, ws_bytes, stream, C.byref(n), C.byref(u)))
    return n.value, u.value


def dev_mergejoin(d_r: int, nbr: int, d_s: int, nbs: int, field, d_ur: int, d_us: int, d_out: int, d_ws: int,
                  ws_bytes: int, stream: int = 0) -> dict:
    res = (C.c_uint64 * 4)()
    check(lib().dbt_dev_mergejoin(d_r, nbr, d_s, nbs, _fld(field), d_ur, d_us, d_out, d_ws, ws_bytes, stream, res))
    return {"nres": res[0], "nunique_R": res[1], "nunique_S": res[2], "later_reads": res[3]}


def dev_hashjoin(d_r: int, nbr: int, d_s: int, nbs: int, field, d_out: int, out_cap_blocks: int, d_ws: int,
                 ws_bytes: int, stream: int = 0) -> int:
    n = C.c_uint64()
    check(lib().dbt_dev_hashjoin(d_r, nbr, d_s, nbs, _fld(field), d_out, out_cap_blocks, d_ws, ws_bytes, stream,
                                 C.byref(n)))
    return n.value


# ---- multi-GPU operators (csrc/dist.cu): one rank per GPU, collective calls -----------------------------
class Dist:
    """One rank of a multi-GPU group (C++ layer: shared-memory control block, peer-memory exchange, no NCCL on the
    data path).  `session` must be the same string on every rank and unique per group (e.g. MASTER_PORT + a nonce)."""

    def __init__(self, session: str, rank: int, world: int, device: int):
        self.L = lib()
        h = C.c_void_p()
        check(self.L.dbt_dist_init(session.encode(), rank, world, device, C.byref(h)))
        self.h, self.rank, self.world = h, rank, world

    def close(self):
        if self.h:
            self.L.dbt_dist_destroy(self.h)
            self.h = None

    def barrier(self):
        check(self.L.dbt_dist_barrier(self.h))

    def trim(self):
        check(self.L.dbt_dist_trim(self.h))

    def set_sub_ranges(self, q: int):
        check(self.L.dbt_dist_set_sub_ranges(self.h, q))

    def sort(self, d_in: int, nblocks: int, field, dedup: bool, d_out: int, cap_blocks: int, stream: int = 0):
        n, m = C.c_uint64(), C.c_uint64()
        check(self.L.dbt_dist_sort(self.h, d_in, nblocks, _fld(field), 1 if dedup else 0, d_out, cap_blocks, stream, C.byref(n),
                                   C.byref(m)))
        return n.value, m.value

    def hashjoin(self, d_r: int, nbr: int, d_s: int, nbs: int, field, d_out: int, cap_blocks: int, stream: int = 0) -> int:
        n = C.c_uint64()
        check(self.L.dbt_dist_hashjoin(self.h, d_r, nbr, d_s, nbs, _fld(field), d_out, cap_blocks, stream, C.byref(n)))
        return n.value

    def mergejoin(self, d_r: int, nbr: int, d_s: int, nbs: int, field, d_out: int, cap_blocks: int, stream: int = 0) -> dict:
        res = (C.c_uint64 * 4)()
        check(self.L.dbt_dist_mergejoin(self.h, d_r, nbr, d_s, nbs, _fld(field), d_out, cap_blocks, stream, res))
        return {"nres": res[0], "nunique_R": res[1], "nunique_S": res[2], "later_reads": res[3]}

    def stats(self) -> dict:
        out = (C.c_double * 16)()
        check(self.L.dbt_dist_stats(self.h, out))
        return {"nvlink_ms": out[0], "bytes_remote": out[1], "bytes_total": out[2], "sub_ranges": int(out[3]),
                "timeline_ms": {"keys_extracted": round(out[4], 3), "splitters": round(out[5], 3), "pushes_enqueued": round(out[6], 3),
                                "first_sub_range_done": round(out[7], 3), "done": round(out[8], 3)}}
