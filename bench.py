#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native tuple operators.

Workload (N=1): BASELINE.json configs[1] -- EliminateDuplicates field=num over 100M records with 10%
duplicate rows (90M distinct 32-bit keys), one B200.  A "step" is one whole operator pass over the
14.016 GB block image: headers -> key extraction -> 4-pass onesweep radix sort of (num, row) pairs ->
adjacent-difference unique -> one record gather into the output image.

  value  records/s, input image already resident in HBM, timed with CUDA events on the operator's
         stream (dbt_dev_dedup through the C-ABI);
  e2e    the same operator through the host-buffer C-ABI call (dbt_host_dedup): pinned host image in,
         pinned host image out, H2D and D2H copies inside the timed region;
  roofline  the dominant kernel's algorithmic bytes / its CUDA-event time, against MEASURED_PEAKS.json;
  cpu_baseline  the untouched reference (oracle/_ref/ref_runner) on a bounded sample, 1 thread.

`--impl reference` times the reference's own CPU implementation instead (rank 0 only).
N>1 (torchrun): every rank holds its own 100M-row shard of one global relation; the C++ multi-GPU layer
(csrc/dist.cu, dbt_dist_*) range-partitions it with sample-sort splitters and moves per-owner block images over NVLink
with the copy engines, key sub-range by key sub-range; every result is checked in the run (bench_verify.py) and the
BASELINE configs[2..4] run as verified legs.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCK_BYTES = 14016
RPB = 100
FIELD = "1"
NMEM = 64


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the untouched reference through oracle/_ref/ref_runner
# ---------------------------------------------------------------------------------------------
def ref_dedup_sample(sample_rows: int, seed: int = 42):
    """One timed run of the reference EliminateDuplicates on a sample with the bench distribution."""
    from oracle import pyoracle as orc  # baseline leg: the one place bench.py may execute oracle/

    orc.build()
    if not orc.ref_available():
        return None
    blocks = orc.gen_syn(seed, sample_rows, (sample_rows * 9) // 10, 0)
    info, out, _ = orc.run_ref("dedup", FIELD, NMEM, blocks)
    return {"seconds": info["seconds"], "rows": sample_rows, "nunique": info["a"]}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    rows = args.ref_rows
    times = []
    for i in range(args.warmup + args.steps):
        r = ref_dedup_sample(rows)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_runner not built"}))
            return 0
        if i >= args.warmup:
            times.append(r["seconds"])
    sec = sum(times) / len(times)
    value = rows / sec
    sample = f"EliminateDuplicates field=num nmem_blocks={NMEM} on a {rows}-record sample of the same distribution (10% duplicate rows), files in tmpfs"
    line = {
        "impl": "reference", "metric": "records_per_second", "value": value, "unit": "records/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "records/s", "cores": 1, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, world):
    return {
        "workload": f"EliminateDuplicates field=num, {args.rows} records per GPU, 10% duplicate rows (BASELINE configs[1]: 100M records, 1 B200)"
                    + (f"; reference arm: timed on a {args.ref_rows}-record sample of the same distribution" if getattr(args, "impl", "") == "reference" else ""),
        "rows_per_gpu": args.rows, "distinct_keys_per_gpu": (args.rows * 9) // 10, "nmem_blocks": NMEM,
        "key_bits": 32, "record_bytes": 140, "cache": "input image (14 GB at 100M rows) is far larger than the 126 MB L2; no flush needed",
        "parallelism": "1 GPU" if world == 1 else f"{world} GPUs, sample-sort key ranges, one exchange of per-owner block images over NVLink (copy engines), pipelined by key sub-range",
    }


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=100_000_000, help="records per GPU")
    ap.add_argument("--ref-rows", type=int, default=4_000_000, help="sample size of the CPU reference runs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--leg-rows", type=int, default=0, help="N>1 legs: records per GPU of the configs[2]/[4] legs (default 125M)")
    ap.add_argument("--join-r", type=int, default=0, help="N>1 legs: total R records of the configs[3] leg (default 100M)")
    ap.add_argument("--join-s", type=int, default=0, help="N>1 legs: total S records of the configs[3] leg (default 1B)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.warmup < 3:
        args.warmup = 3  # timing rule: at least 3 warm-up steps

    import torch

    dbt = importlib.import_module("database-technology-algorithms_b200")
    L = dbt.lib()
    if not torch.cuda.is_available() or L.dbt_device_count() == 0:
        raise SystemExit("bench.py: no CUDA device visible -- the product has no CPU fallback")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
        return bench_dist(args, dbt, rank, world, local_rank)

    peak, peak_src = load_peaks()
    n = args.rows
    U = (n * 9) // 10
    nblocks = (n + RPB - 1) // RPB
    img_bytes = nblocks * BLOCK_BYTES
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    d_in = torch.empty(img_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(img_bytes, dtype=torch.uint8, device=dev)
    wsb = dbt.dev_ws_bytes(dbt.OP_DEDUP, nblocks, 0, FIELD)
    d_ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    dbt.check(L.dbt_gen_syn(42, n, U, 0, 0, n, 0, d_in.data_ptr(), sp))
    torch.cuda.synchronize()

    def step():
        return dbt.dev_dedup(d_in.data_ptr(), nblocks, FIELD, d_out.data_ptr(), d_ws.data_ptr(), wsb, sp)

    for _ in range(args.warmup):
        rows, uniq = step()
    assert rows == n and uniq == U, f"wrong result: rows={rows} nunique={uniq}, expected {n}/{U}"

    # ---- timed region: exactly K steps, CUDA events on the operator's stream ------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    L.dbt_stage_timing_enable(1)
    L.dbt_stage_timing_reset()
    launches0 = L.dbt_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    launches = int(L.dbt_kernel_launches() - launches0)
    stages = dbt.stage_report()
    L.dbt_stage_timing_enable(0)
    ms_step = ms_total / args.steps
    value = n / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (live CUDA-event stage timing) -------------------------
    alg_bytes = {  # algorithmic bytes per launch (DESIGN.md "Kernels")
        "record_gather": 280.0 * U,       # 140 B read + 140 B written per surviving record
        "onesweep_pass": 16.0 * n,        # 8 B read + 8 B written per (key,row) pair
        "extract_keys": 12.0 * n,         # 4 B key read (sector-granular in practice) + key + recid columns written
        "unique": 8.0 * n + 4.0 * U,      # (the digit histogram is made inside the extraction: no key read of its own)
    }
    dom = max(stages.items(), key=lambda kv: kv[1][0])[0] if stages else None
    roofline = None
    if dom in alg_bytes:
        ms_dom, _ = stages[dom]
        kernel_launches = {"onesweep_pass": 4 * args.steps}.get(dom, args.steps)
        avg_ms = ms_dom / kernel_launches
        achieved = alg_bytes[dom] / (avg_ms * 1e-3) / 1e9
        roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": load_traffic(dom), "peak_source": peak_src,
                    "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": alg_bytes[dom]}
    stage_ms = {k: round(v[0] / args.steps, 4) for k, v in stages.items()}
    per_stage_frac = {}
    for k, b in alg_bytes.items():
        if k in stages and stages[k][0] > 0:
            launches_k = {"onesweep_pass": 4}.get(k, 1)
            per_stage_frac[k] = round(b * launches_k / (stages[k][0] / args.steps * 1e-3) / 1e9 / peak, 4)

    # ---- e2e: host buffers through dbt_host_dedup, copies inside the timed region ----------------
    e2e = None
    if not args.no_e2e:
        import ctypes as C

        h_in, h_out = C.c_void_p(), C.c_void_p()
        dbt.check(L.dbt_host_alloc(C.byref(h_in), img_bytes))
        dbt.check(L.dbt_host_alloc(C.byref(h_out), img_bytes))
        torch.cuda.synchronize()
        # fill the pinned host image from the device copy (setup, untimed)
        host_t = torch.frombuffer((C.c_uint8 * img_bytes).from_address(h_in.value), dtype=torch.uint8)
        host_t.copy_(d_in)
        nr, nu = C.c_uint64(), C.c_uint64()
        out_bytes = ((U + RPB - 1) // RPB) * BLOCK_BYTES

        def e2e_step():
            dbt.check(L.dbt_host_dedup(h_in, nblocks, ord(FIELD), h_out, local_rank, C.byref(nr), C.byref(nu)))

        for _ in range(2):
            e2e_step()
        assert nu.value == U
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        sec_sync = (time.perf_counter() - t0) / args.steps
        # The same steps as a stream of jobs, two in flight (dbt_host_dedup_begin / dbt_host_job_wait): step i's
        # download overlaps step i+1's upload.  Every step still uploads its input and downloads its result inside
        # the timed region, which ends when the last result is in host memory.
        h_out2 = C.c_void_p()
        sec = sec_sync
        api = "dbt_host_dedup (C-ABI, pinned host image in/out), one synchronous call per step"
        try:
            dbt.check(L.dbt_host_alloc(C.byref(h_out2), out_bytes))
        except dbt.DbtError:
            h_out2 = None  # not enough pinned host memory for a second result buffer: keep the synchronous figure
        if h_out2 is not None:
            outs = [h_out, h_out2]
            res4 = (C.c_uint64 * 4)()

            def pipelined(steps):
                pending = [False, False]
                for i in range(steps):
                    sl = i % 2
                    if pending[sl]:
                        dbt.check(L.dbt_host_job_wait(sl, res4))
                        assert res4[1] == U
                    dbt.check(L.dbt_host_dedup_begin(sl, h_in, nblocks, ord(FIELD), outs[sl], local_rank))
                    pending[sl] = True
                for sl in ((steps % 2), 1 - (steps % 2)):  # oldest first
                    if pending[sl]:
                        dbt.check(L.dbt_host_job_wait(sl, res4))
                        assert res4[1] == U

            pipelined(3)
            t0 = time.perf_counter()
            pipelined(args.steps)
            sec = (time.perf_counter() - t0) / args.steps
            api = ("dbt_host_dedup_begin / dbt_host_job_wait (C-ABI, pinned host image in/out), two jobs in flight: "
                   "step i's download overlaps step i+1's upload; timed until the last result is in host memory")
            out2_t = torch.frombuffer((C.c_uint8 * out_bytes).from_address(h_out2.value), dtype=torch.uint8)
            out1_t = torch.frombuffer((C.c_uint8 * out_bytes).from_address(h_out.value), dtype=torch.uint8)
            assert torch.equal(out1_t[:: 4099], out2_t[:: 4099])  # both slots landed the same image
            del out1_t, out2_t
            L.dbt_host_free(h_out2)
        e2e = {"value": n / sec, "unit": "records/s", "h2d_bytes_per_step": img_bytes, "d2h_bytes_per_step": out_bytes,
               "ms_per_step": sec * 1e3, "api": api,
               "synchronous_call": {"value": n / sec_sync, "ms_per_step": sec_sync * 1e3,
                                    "api": "dbt_host_dedup, one blocking call per step"}}
        # spot check of the host result: first block header + record count
        out_t = torch.frombuffer((C.c_uint8 * out_bytes).from_address(h_out.value), dtype=torch.uint8)
        assert int(out_t[4:8].view(torch.int32)[0]) == 100
        del host_t, out_t
        L.dbt_host_free(h_in)
        L.dbt_host_free(h_out)
        dbt.check(L.dbt_host_trim())  # give the slots' cached device buffers back before the extras

    clocks = sampler.stop()  # sampled across both timed regions (device scope and e2e)

    # ---- cpu baseline: the reference itself on a bounded sample ------------------------------------
    cpu = None
    if not args.no_cpu:
        r = ref_dedup_sample(args.ref_rows)
        if r is not None:
            cpu = {"value": r["rows"] / r["seconds"], "unit": "records/s", "cores": 1, "kind": "reference",
                   "sample": f"reference EliminateDuplicates (oracle/_ref/ref_runner, single-threaded as shipped) on a "
                             f"{r['rows']}-record sample of the same distribution, nmem_blocks={NMEM}, files in tmpfs; "
                             f"host has {os.cpu_count()} cores"}

    extra = {}
    if not args.no_extra:
        del d_out, d_ws, d_in
        torch.cuda.empty_cache()
        extra = extra_measurements(dbt, torch, dev, peak)

    line = {
        "metric": "records_per_second", "value": value, "unit": "records/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args, 1),
        "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "gpu_launches": launches,
        "stage_ms_per_step": stage_ms, "stage_hbm_frac": per_stage_frac, "extra": extra,
    }
    summary = {"cfg1_dedup_100M": {"ms": round(ms_step, 2), "Grec_s": round(value / 1e9, 2), "e2e_ms": round(e2e["ms_per_step"], 1) if e2e else None}}
    for k, v in extra.items():  # compact: the driver keeps the last 1500 characters of stdout
        if isinstance(v, dict):
            e = v.get("dbt_b200_entry_point", v)
            summary[k] = {a: (round(b, 2) if isinstance(b, float) else b) for a, b in e.items()
                          if a in ("ms", "nres", "error", "nunique", "hbm_frac_of_measured_at_68B_per_record", "s_passes_hbm_frac_of_measured")}
            if "reference" in v and v["reference"]:
                summary[k]["ref_cpu_ms"] = round(v["reference"]["ms"], 1)
    line["summary"] = summary
    print(json.dumps(line))
    return 0


def _timed(torch, dist, dev, stream, fn, steps, warmup):
    """W untimed + K timed calls of fn, CUDA events on `stream`, barrier + synchronize on both sides, max over ranks."""
    out = None
    for _ in range(warmup):
        out = fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(steps):
        out = fn()
    e1.record(stream)
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()) / steps, out


def _ok(d: dict) -> bool:
    return all(v for v in d.values() if isinstance(v, bool))


def bench_dist(args, dbt, rank, world, local_rank):
    """N>1, one process per GPU.  Headline: one global relation of N x rows records (10% duplicate rows globally), rank r
    holds rows [r*rows, (r+1)*rows); a step = distributed EliminateDuplicates through the C++ multi-GPU layer
    (dbt_dist_sort: sample-sort splitters, per-owner block images carried over NVLink by the copy engines sub-range by
    sub-range, the single-GPU operator on every landed sub-range); the concatenation of the ranks' outputs is the
    globally sorted duplicate-free file (weak scaling: rows per GPU fixed).  torch.distributed (NCCL) does the contract's
    barriers / max-over-ranks and the reductions of the independent result checks; it moves no record.
    Legs (after the headline, each verified in the run): BASELINE configs[2] MergeSort field=str (125M records per GPU =
    1B at 8), configs[3] HashJoin field=num R=100M x S=1B split over the GPUs (strong scaling), uniform and Zipf(1.1),
    configs[4] MergeJoin field=num+str (62.5M + 62.5M records per GPU = 2 x 500M at 8), and a small instance of every
    operator compared bit for bit with the CPU oracle."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    import bench_verify as V

    L = dbt.lib()
    dev = torch.device("cuda", local_rank)
    tok = [f"{os.environ.get('MASTER_PORT', '0')}_{os.getpid()}_{int(time.time() * 1e3) % 1000000}"]
    dist.broadcast_object_list(tok, src=0)
    d = dbt.Dist(tok[0], rank, world, local_rank)
    peak, peak_src = load_peaks()
    n = args.rows
    n_total = n * world
    U = (n_total * 9) // 10
    nblocks = (n + RPB - 1) // RPB
    img_bytes = nblocks * BLOCK_BYTES
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    steps, warmup = args.steps, max(args.warmup, 3)
    d_in = torch.empty(img_bytes, dtype=torch.uint8, device=dev)
    dbt.check(L.dbt_gen_syn(42, n_total, U, 0, rank * n, n, 0, d_in.data_ptr(), sp))
    cap = nblocks + nblocks // 4 + 64  # splitters come from a sample: leave headroom for an uneven cut
    d_out = torch.empty(cap * BLOCK_BYTES, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def step():
        return d.sort(d_in.data_ptr(), nblocks, FIELD, True, d_out.data_ptr(), cap, sp)

    for _ in range(warmup):
        out_rows, _ = step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        L.dbt_stage_timing_enable(1)
        L.dbt_stage_timing_reset()
    launches0 = L.dbt_kernel_launches()
    ms_step, (out_rows, recv_rows) = _timed(torch, dist, dev, stream, step, steps, 0)
    launches = int(L.dbt_kernel_launches() - launches0)
    st = d.stats()
    stages = dbt.stage_report() if rank == 0 else {}
    if rank == 0:
        L.dbt_stage_timing_enable(0)
    value = n_total / (ms_step * 1e-3)
    x = torch.tensor([st["nvlink_ms"], st["bytes_remote"]], device=dev, dtype=torch.float64)
    dist.all_reduce(x, op=dist.ReduceOp.MAX)
    nv_ms, nv_bytes = float(x[0].item()), float(x[1].item())

    # ---- independent check of the headline result (torch only, full size) ------------------------------
    checks = {}
    try:
        checks["cfg1_dedup_num"] = V.check_dedup_u32(d_in, n, d_out, out_rows, 1, dist)
        checks["cfg1_dedup_num"]["rows_expected"] = U
        checks["cfg1_dedup_num"]["count_ok"] = checks["cfg1_dedup_num"]["rows"] == U
    except Exception as e:  # noqa: BLE001
        checks["cfg1_dedup_num"] = {"error": str(e)[:160]}
    torch.cuda.empty_cache()

    # ---- e2e: pinned host shard in, pinned host result out, every step's copies inside the timed region; the upload of
    # step i+1 and the download of step i-1 run on their own streams under step i's operator ---------------------------
    e2e = None
    if not args.no_e2e:
        e2e = e2e_dist(args, dbt, d, torch, dist, dev, d_in, nblocks, cap, n_total, world, U)
    clocks = sampler.stop() if rank == 0 else None
    del d_in, d_out
    torch.cuda.empty_cache()

    legs = {}
    if not args.no_extra:
        legs = dist_legs(args, dbt, d, torch, dist, V, dev, rank, world, checks)

    all_ok = all(_ok(c) for c in checks.values())  # a comparison that came out wrong fails the run; a check that could not run is reported
    if rank == 0:
        dom = max(stages.items(), key=lambda kv: kv[1][0])[0] if stages else None
        roofline = None
        if dom == "record_gather":
            # per step on this rank: the per-owner gathers of its n rows + the owner-side gathers of the rows it emits
            ms_dom = stages[dom][0] / steps
            rows_moved = n + out_rows
            achieved = 280.0 * rows_moved / (ms_dom * 1e-3) / 1e9
            roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": None, "traffic_note": "no ncu capture of the multi-GPU step (ncu is single-GPU only here)",
                        "peak_source": peak_src, "ms_per_step_all_launches": ms_dom, "algorithmic_bytes_per_step": 280.0 * rows_moved}
        nvlink = {"what": "record bytes this GPU sends to the other GPUs per step (copy engines, per-owner block images) and the "
                          "duration of that phase, first copy .. last flag (max over ranks)",
                  "bytes_sent_per_gpu": nv_bytes, "phase_ms": nv_ms, "sub_ranges": st["sub_ranges"],
                  "achieved_gbs_per_direction": nv_bytes / (nv_ms * 1e-3) / 1e9 if nv_ms > 0 else None,
                  "peak_gbs_per_direction": 780.0, "peak_source": "profiles/micro/p2p_scatter.cu: cudaMemcpyPeer between two B200 of this pool"}
        summary = {"checks_ok": all_ok, "check_errors": sum(1 for c in checks.values() if "error" in c),
                   "cfg1": {"ms": round(ms_step, 2), "Grec_s": round(value / 1e9, 2), "ok": _ok(checks.get("cfg1_dedup_num", {}))}}
        for k, v in legs.items():
            summary[k] = {a: b for a, b in v.items() if a in ("ms", "Grec_s", "Gtup_s", "ok", "rows", "nres", "error", "skipped")}
        line = {
            "metric": "records_per_second", "value": value, "unit": "records/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": workload_config(args, world),
            "e2e": e2e, "roofline": roofline, "cpu_baseline": None, "clocks": clocks, "gpu_launches": launches,
            "nvlink": nvlink, "stage_ms_per_step": {k: round(v[0] / steps, 4) for k, v in stages.items()},
            "timeline_ms_last_step_rank0": st["timeline_ms"], "checks": checks, "legs": legs, "summary": summary,
        }
        print(json.dumps(line))
    d.barrier()
    d.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0 if all_ok else 1


def e2e_dist(args, dbt, d, torch, dist, dev, d_in, nblocks, cap, n_total, world, U):
    """Every rank: pinned host shard -> device -> distributed dedup -> pinned host result, all inside the timed region.
    Two device input buffers and two output buffers: step i+1's upload and step i-1's download run on their own streams
    while step i's operator runs (PCIe is full duplex; the operator itself takes ~25 ms of a ~300 ms step)."""
    img_bytes = nblocks * BLOCK_BYTES
    try:
        import psutil

        avail = psutil.virtual_memory().available
    except Exception:  # noqa: BLE001
        avail = 0
    out_bytes = cap * BLOCK_BYTES
    host_out_bytes = (nblocks + nblocks // 16 + 8) * BLOCK_BYTES  # a rank emits about its share (0.9 of its rows here)
    need = (img_bytes + host_out_bytes) * world * 1.15
    free_dev, _ = torch.cuda.mem_get_info(dev)
    if avail < need or free_dev < 2 * img_bytes + 2 * out_bytes + (8 << 30):
        return {"value": None, "unit": "records/s", "h2d_bytes_per_step": img_bytes * world, "d2h_bytes_per_step": None,
                "skipped": f"needs {need / 1e9:.0f} GB of pinned host memory ({avail / 1e9:.0f} available) and "
                           f"{(2 * img_bytes + 2 * out_bytes) / 1e9:.0f} GB of device memory ({free_dev / 1e9:.0f} free)"}
    h_in = torch.empty(img_bytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(host_out_bytes, dtype=torch.uint8).pin_memory()
    h_in.copy_(d_in)
    bufs_in = [d_in, torch.empty_like(d_in)]
    bufs_out = [torch.empty(out_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    s_up, s_down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()
    sp = main.cuda_stream

    def run(steps):
        up = [None, None]
        down = [None, None]
        with torch.cuda.stream(s_up):
            bufs_in[0].copy_(h_in, non_blocking=True)
            up[0] = s_up.record_event()
        out_b = 0
        for i in range(steps):
            k = i & 1
            if i + 1 < steps:  # next step's input comes in under this step's operator
                if down[k ^ 1] is not None:
                    pass  # (its input buffer was consumed by step i-1, which has returned)
                with torch.cuda.stream(s_up):
                    bufs_in[k ^ 1].copy_(h_in, non_blocking=True)
                    up[k ^ 1] = s_up.record_event()
            main.wait_event(up[k])
            if down[k] is not None:
                main.wait_event(down[k])  # the output buffer of step i-2 is home
            rows, _ = d.sort(bufs_in[k].data_ptr(), nblocks, FIELD, True, bufs_out[k].data_ptr(), cap, sp)  # returns when done
            assert rows > 0
            nb_out = (rows + RPB - 1) // RPB
            out_b = nb_out * BLOCK_BYTES
            assert out_b <= host_out_bytes
            with torch.cuda.stream(s_down):  # (one host buffer: the downloads are ordered on their stream)
                h_out[:out_b].copy_(bufs_out[k][:out_b], non_blocking=True)
                down[k] = s_down.record_event()
        torch.cuda.synchronize()
        return out_b

    run(2)
    dist.barrier()
    t0 = time.perf_counter()
    ob = run(args.steps)
    dist.barrier()
    sec = torch.tensor([(time.perf_counter() - t0) / args.steps], device=dev, dtype=torch.float64)
    dist.all_reduce(sec, op=dist.ReduceOp.MAX)
    tot = torch.tensor([ob], device=dev, dtype=torch.float64)
    dist.all_reduce(tot)
    del h_in, h_out, bufs_out
    return {"value": n_total / float(sec.item()), "unit": "records/s", "h2d_bytes_per_step": img_bytes * world,
            "d2h_bytes_per_step": float(tot.item()), "ms_per_step": float(sec.item()) * 1e3,
            "api": "per rank: pinned host shard -> cudaMemcpyAsync -> dbt_dist_sort (C-ABI, dedup) -> cudaMemcpyAsync -> pinned host "
                   "result; uploads and downloads of neighbouring steps overlap the operator on their own streams"}


def dist_legs(args, dbt, d, torch, dist, V, dev, rank, world, checks):
    """BASELINE configs[2..4] at the scale the GPU count allows, each timed (3 steps after 1 warm-up, CUDA events, max
    over ranks) and checked in the run, plus a small instance of every operator against the CPU oracle."""
    import ctypes as C

    L = dbt.lib()
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    legs = {}

    def gen(seed, n_total, U, kind, row0, nrows):
        nb = (nrows + RPB - 1) // RPB
        t = torch.empty(nb * BLOCK_BYTES, dtype=torch.uint8, device=dev)
        dbt.check(L.dbt_gen_syn(seed, n_total, U, kind, row0, nrows, 0, t.data_ptr(), sp))
        torch.cuda.synchronize()
        return t, nb

    def agree(local_error):
        """A leg that failed on one rank must be abandoned by all of them (the operators are collective)."""
        f = torch.tensor([1 if local_error else 0], device=dev)
        dist.all_reduce(f, op=dist.ReduceOp.MAX)
        return bool(f.item())

    # ---- configs[2]: MergeSort field=str, 125M records per GPU (1B at 8 GPUs) --------------------------------------
    try:
        d.trim()
        n2 = args.leg_rows or 125_000_000
        img, nb = gen(77, n2 * world, n2 * world, 1, rank * n2, n2)
        cap = nb + nb // 4 + 64
        out = torch.empty(cap * BLOCK_BYTES, dtype=torch.uint8, device=dev)
        ms, (rows, _) = _timed(torch, dist, dev, stream, lambda: d.sort(img.data_ptr(), nb, "2", False, out.data_ptr(), cap, sp), 3, 1)
        st = d.stats()
        chk = V.check_sort(img, n2, out, rows, V.str_key64, dist)
        checks["cfg2_sort_str"] = chk
        legs["cfg2_sort_str"] = {"workload": f"MergeSort field=str, {n2} records per GPU x {world} GPUs = {n2 * world} records (BASELINE configs[2]: 1B on 8)",
                                 "ms": round(ms, 2), "Grec_s": round(n2 * world / ms / 1e6, 2), "rows": chk["rows"], "ok": _ok(chk),
                                 "nvlink_gbs_per_direction": round(st["bytes_remote"] / max(st["nvlink_ms"], 1e-9) / 1e6, 1), "timeline_ms": st["timeline_ms"]}
        del img, out
    except Exception as e:  # noqa: BLE001
        legs["cfg2_sort_str"] = {"error": str(e)[:200]}
    torch.cuda.empty_cache()

    # ---- north_star: MergeSort field=num at operator scope, 125M records per GPU (1B records on 8 GPUs) ---------------
    try:
        d.trim()
        n2 = args.leg_rows or 125_000_000
        img, nb = gen(78, n2 * world, 1 << 32, 1, rank * n2, n2)  # num uniform over the full 32-bit range
        cap = nb + nb // 4 + 64
        out = torch.empty(cap * BLOCK_BYTES, dtype=torch.uint8, device=dev)
        ms, (rows, _) = _timed(torch, dist, dev, stream, lambda: d.sort(img.data_ptr(), nb, "1", False, out.data_ptr(), cap, sp), 3, 1)
        st = d.stats()
        chk = V.check_sort(img, n2, out, rows, lambda im, m: V.column(im, m, 1), dist)
        checks["northstar_sort_num"] = chk
        legs["northstar_sort_num"] = {"workload": f"MergeSort field=num, {n2} records per GPU x {world} GPUs = {n2 * world} records, full 32-bit keys (north_star: 1B-record num-key MergeSort; 140 GB of records only fit on 8 GPUs)",
                                      "ms": round(ms, 2), "Grec_s": round(n2 * world / ms / 1e6, 2), "rows": chk["rows"], "ok": _ok(chk),
                                      "nvlink_gbs_per_direction": round(st["bytes_remote"] / max(st["nvlink_ms"], 1e-9) / 1e6, 1), "timeline_ms": st["timeline_ms"]}
        del img, out
    except Exception as e:  # noqa: BLE001
        legs["northstar_sort_num"] = {"error": str(e)[:200]}
    torch.cuda.empty_cache()

    # ---- configs[3]: HashJoin field=num, R = 100M x S = 1B over the GPUs (strong scaling), uniform and Zipf(1.1) -------
    for kind, label in ((1, "uniform"), (4, "zipf1.1")):
        name = f"cfg3_hashjoin_{label}"
        try:
            d.trim()
            NR, NS, D = args.join_r or 100_000_000, args.join_s or 1_000_000_000, 100_000_000
            nr, ns = NR // world // RPB * RPB, NS // world // RPB * RPB
            r_img, nbr = gen(7, NR, D, 1, rank * nr, nr)
            s_img, nbs = gen(9, NS, D, kind, rank * ns, ns)
            cap = int(nbs * 0.78) + 64
            out = torch.empty(cap * BLOCK_BYTES, dtype=torch.uint8, device=dev)
            ms, k = _timed(torch, dist, dev, stream,
                           lambda: d.hashjoin(r_img.data_ptr(), nbr, s_img.data_ptr(), nbs, "1", out.data_ptr(), cap, sp), 3, 1)
            chk = V.check_semijoin_u32(r_img, nr, s_img, ns, out, k, 1, D, dist)
            checks[name] = chk
            legs[name] = {"workload": f"HashJoin field=num, R={nr * world} x S={ns * world} records over {world} GPUs, S keys {label} over [0,1e8) "
                                      "(BASELINE configs[3]); semi-join: S rows in S order whose key is in keys(R)",
                          "ms": round(ms, 2), "Gtup_s": round(ns * world / ms / 1e6, 2), "nres": chk["rows"], "ok": _ok(chk)}
            del r_img, s_img, out
        except Exception as e:  # noqa: BLE001
            legs[name] = {"error": str(e)[:200]}
        torch.cuda.empty_cache()

    # ---- configs[4]: MergeJoin field=num+str, 62.5M + 62.5M records per GPU (2 x 500M at 8 GPUs) ---------------------
    try:
        d.trim()
        n4 = args.leg_rows // 2 if args.leg_rows else 62_500_000
        U4 = max(1000, int(0.3 * n4 * world))
        r_img, nb = gen(21, n4 * world, U4, 1, rank * n4, n4)
        s_img, _ = gen(21 ^ 0x5EED, n4 * world, U4, 3, rank * n4, n4)
        cap = nb + nb // 4 + 64
        out = torch.empty(cap * BLOCK_BYTES, dtype=torch.uint8, device=dev)
        ms, info = _timed(torch, dist, dev, stream,
                          lambda: d.mergejoin(r_img.data_ptr(), nb, s_img.data_ptr(), nb, "3", out.data_ptr(), cap, sp), 3, 1)
        chk = V.check_mergejoin_composite(r_img, n4, s_img, n4, out, info["nres"], dist)
        checks["cfg4_mergejoin_numstr"] = chk
        legs["cfg4_mergejoin_numstr"] = {"workload": f"MergeJoin field=num+str, 2 x {n4 * world} records over {world} GPUs (BASELINE configs[4]: 2 x 500M on 8)",
                                         "ms": round(ms, 2), "Gtup_s": round(2 * n4 * world / ms / 1e6, 2), "nres": chk["rows"], "ok": _ok(chk)}
        del r_img, s_img, out
    except Exception as e:  # noqa: BLE001
        legs["cfg4_mergejoin_numstr"] = {"error": str(e)[:200]}
    torch.cuda.empty_cache()

    # ---- a small instance of every operator, bit for bit against the CPU oracle (rank 0 runs the oracle) -----------
    try:
        d.trim()
        from oracle import pyoracle as orc  # checker only

        import numpy as np

        nloc = 20_000
        nbl = nloc // RPB
        r_img, _ = gen(5, nloc * world, nloc * world // 3, 1, rank * nloc, nloc)
        s_img, _ = gen(5 ^ 0x5EED, nloc * world, nloc * world // 3, 3, rank * nloc, nloc)
        cap = nbl * world + 8
        out = torch.empty(cap * BLOCK_BYTES, dtype=torch.uint8, device=dev)
        got = {}

        def rows_of(k):
            nbk = (k + RPB - 1) // RPB
            return out[: nbk * BLOCK_BYTES].cpu().numpy().tobytes(), k

        k, _ = d.sort(r_img.data_ptr(), nbl, "3", False, out.data_ptr(), cap, sp)
        got["sort3"] = rows_of(k)
        k, _ = d.sort(r_img.data_ptr(), nbl, "1", True, out.data_ptr(), cap, sp)
        got["dedup1"] = rows_of(k)
        k = d.hashjoin(r_img.data_ptr(), nbl, s_img.data_ptr(), nbl, "1", out.data_ptr(), cap, sp)
        got["hashjoin1"] = rows_of(k)
        k = d.mergejoin(r_img.data_ptr(), nbl, s_img.data_ptr(), nbl, "3", out.data_ptr(), cap, sp)["nres"]
        got["mergejoin3"] = rows_of(k)
        every = [None] * world
        dist.all_gather_object(every, got)
        res = {}
        if rank == 0:
            orc.build()
            R = orc.gen_syn(5, nloc * world, nloc * world // 3, 1)
            S = orc.gen_syn(5 ^ 0x5EED, nloc * world, nloc * world // 3, 3)
            want = {"sort3": orc.sort(R, "3"), "dedup1": orc.dedup(R, "1"), "hashjoin1": orc.hashjoin(R, S, "1"),
                    "mergejoin3": orc.mergejoin(R, S, "3")[0]}
            for name, w in want.items():
                rows = [orc.rows_of(orc.as_blocks(np.frombuffer(g[name][0], dtype=np.uint8).copy()))[: g[name][1]] for g in every]
                res[name] = np.concatenate(rows).tobytes() == orc.rows_of(w).tobytes()  # every row, all 140 bytes, global order
        checks["oracle_small"] = res if rank == 0 else {"rank": True}
        legs["oracle_small"] = {"workload": f"{nloc} records per GPU: sort field 3, dedup field 1, hashjoin field 1, mergejoin field 3 vs oracle/dbt_oracle.c, all 140 bytes of every row",
                                "ok": _ok(res) if rank == 0 else True}
        del r_img, s_img, out
    except Exception as e:  # noqa: BLE001
        legs["oracle_small"] = {"error": str(e)[:200]}
    torch.cuda.empty_cache()
    return legs


def load_traffic(kernel: str):
    """dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(kernel)
    return None


def extra_measurements(dbt, torch, dev, peak):
    """north_star side targets, reported next to the headline (not the headline itself)."""
    import ctypes as C

    L = dbt.lib()
    out = {}
    sp = torch.cuda.current_stream().cuda_stream
    # 1B-record num-key pair sort on one B200 (records cannot be resident: 140 GB)
    try:
        n = 1_000_000_000
        g = torch.Generator(device=dev).manual_seed(1)
        keys = torch.randint(-2**31, 2**31, (n,), dtype=torch.int32, device=dev, generator=g)
        k1, k2 = torch.empty_like(keys), torch.empty_like(keys)
        v1, v2 = torch.empty_like(keys), torch.empty_like(keys)
        wsb = L.dbt_sort_pairs_ws_bytes(n)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        alt = C.c_int()
        times = []
        for it in range(4):
            k1.copy_(keys)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dbt.check(L.dbt_sort_pairs_u32(k1.data_ptr(), k2.data_ptr(), v1.data_ptr(), v2.data_ptr(), n, 0, 32,
                                           ws.data_ptr(), wsb, sp, C.byref(alt)))
            e1.record()
            torch.cuda.synchronize()
            if it:
                times.append(e0.elapsed_time(e1))
        ms = sum(times) / len(times)
        ko = k2 if alt.value else k1
        ok = True
        for lo in range(0, n, 100_000_000):  # sortedness as unsigned 32-bit, chunk by chunk
            c = ko[lo:min(n, lo + 100_000_001)].to(torch.int64) & 0xFFFFFFFF
            ok = ok and bool((c[1:] >= c[:-1]).all().item())
        assert ok, "1B pair sort produced an unsorted result"
        out["pair_sort_1B_u32"] = {"records_per_s": n / (ms * 1e-3), "ms": ms,
                                   "hbm_frac_of_measured_at_68B_per_record": 68.0 * n / (ms * 1e-3) / 1e9 / peak,
                                   "note": "keys+values already in HBM; one read of the keys for OR/AND + the four byte histograms (4 B) and 4 passes x 16 B = 68 B per pair (SURVEY.md 8d)"}
        del keys, k1, k2, v1, v2, ws
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out["pair_sort_1B_u32"] = {"error": str(e)[:200]}
    # BASELINE configs[0]: MergeSort field=num on a 1M-record block file, nmem_blocks=64, through the real
    # file-based drop-in entry point (as main.o would call it), files in tmpfs; the reference binary beside it
    try:
        import shutil
        import tempfile

        n0 = 1_000_000
        nb0 = n0 // RPB
        img = torch.empty(nb0 * BLOCK_BYTES, dtype=torch.uint8, device=dev)
        dbt.check(L.dbt_gen_syn(42, n0, 300_000, 1, 0, n0, 0, img.data_ptr(), sp))  # ~3.3 rows per key, like main.cpp:47
        torch.cuda.synchronize()
        root = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
        d = tempfile.mkdtemp(prefix="dbtcfg0_", dir=root)
        cwd = os.getcwd()
        try:
            img.cpu().numpy().tofile(os.path.join(d, "file.bin"))
            os.chdir(d)
            f = getattr(L, dbt.CXX_ENTRY_POINTS["MergeSort"])
            f.restype = None
            a, b, c = C.c_uint(), C.c_uint(), C.c_uint()
            name = C.create_string_buffer(64)
            best = 1e9
            for _ in range(4):
                t0 = time.perf_counter()
                f(b"file.bin", C.c_ubyte(ord(FIELD)), None, C.c_uint(NMEM), name, C.byref(a), C.byref(b), C.byref(c))
                best = min(best, time.perf_counter() - t0)
            entry = {"ms": best * 1e3, "records_per_s": n0 / best, "nsorted_segs": a.value, "npasses": b.value, "nios": c.value,
                     "outfile": name.value.decode(), "path": "file -> pinned (parallel pread) -> HBM -> kernels -> pinned -> file"}
            ref = None
            try:  # cpu-baseline leg: the untouched reference on the same file
                from oracle import pyoracle as orc

                if orc.ref_available():
                    p = subprocess.run([orc.REF_RUNNER, "sort", FIELD, str(NMEM), "file.bin"], cwd=d, capture_output=True, text=True)
                    info = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
                    ref = {"ms": info["seconds"] * 1e3, "records_per_s": n0 / info["seconds"], "nsorted_segs": info["a"],
                           "npasses": info["b"], "nios": info["nios"], "cores": 1}
            except Exception:  # noqa: BLE001
                ref = None
            out["cfg0_mergesort_1M_record_file"] = {"dbt_b200_entry_point": entry, "reference": ref}
        finally:
            os.chdir(cwd)
            shutil.rmtree(d, ignore_errors=True)
        del img
    except Exception as e:  # noqa: BLE001
        out["cfg0_mergesort_1M_record_file"] = {"error": str(e)[:200]}
    # HashJoin field=num, R=100M x S=400M rows (BASELINE configs[3] asks for S=1B: 140 GB of S records do not
    # fit beside R and the output on one 180 GB GPU, so the single-GPU figure uses the largest S that does)
    for kind, label in ((1, "uniform"), (4, "zipf1.1")):
        try:
            nr, ns, D = 100_000_000, 400_000_000, 100_000_000
            nbr, nbs = nr // RPB, ns // RPB
            d_r = torch.empty(nbr * BLOCK_BYTES, dtype=torch.uint8, device=dev)
            d_s = torch.empty(nbs * BLOCK_BYTES, dtype=torch.uint8, device=dev)
            d_o = torch.empty(nbs * BLOCK_BYTES, dtype=torch.uint8, device=dev)
            dbt.check(L.dbt_gen_syn(7, nr, D, 1, 0, nr, 0, d_r.data_ptr(), sp))
            dbt.check(L.dbt_gen_syn(9, ns, D, kind, 0, ns, 0, d_s.data_ptr(), sp))
            wsb = dbt.dev_ws_bytes(dbt.OP_HASHJOIN, nbr, nbs, FIELD)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            L.dbt_stage_timing_enable(1)
            times = []
            for it in range(4):
                torch.cuda.synchronize()
                L.dbt_stage_timing_reset()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                k = dbt.dev_hashjoin(d_r.data_ptr(), nbr, d_s.data_ptr(), nbs, FIELD, d_o.data_ptr(), nbs, ws.data_ptr(), wsb, sp)
                e1.record()
                torch.cuda.synchronize()
                if it:
                    times.append(e0.elapsed_time(e1))
            rep = dbt.stage_report()
            L.dbt_stage_timing_enable(0)
            ms = sum(times) / len(times)
            probe_ms = rep["hash_probe"][0] + rep.get("record_gather", (0.0, 0))[0]  # the two streaming passes over S
            sel = k / ns
            out[f"hashjoin_R100M_S400M_{label}"] = {
                "probe_tuples_per_s": ns / (ms * 1e-3), "ms": ms, "nres": k, "selectivity": sel,
                "stage_ms": {a: round(b[0], 3) for a, b in rep.items()},
                "s_passes_hbm_frac_of_measured": (140 + 140 * sel) * ns / (probe_ms * 1e-3) / 1e9 / peak,
                "s_passes_actual_bytes_frac_of_measured": (68 + 140 + 140 * sel) * ns / (probe_ms * 1e-3) / 1e9 / peak,
                "note": "semi-join (reference semantics): S rows in S order whose key is in keys(R).  R: extraction + "
                        "direct-address bitmap (L2-resident: keys(R) span < 2^29).  S: a key pass (8 bytes per row through "
                        "64-byte-fill loads: ~68 B of DRAM reads per row, ncu), the scan of the per-block match counts, and a "
                        "streaming pass over the image that copies the matching records to their final place: algorithmic "
                        "140 B read per S row + 140 B written per match; actual 68 + 140 B read per S row"}
            del d_r, d_s, d_o, ws
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            out[f"hashjoin_R100M_S400M_{label}"] = {"error": str(e)[:200]}
    # Out-of-core mode measured where the in-core figure exists: the headline workload through
    # dbt_host_dedup with the chunk forced to a quarter of the image (4 runs), pinned host image in/out.
    try:
        n1, U1 = 100_000_000, 90_000_000
        nb1 = n1 // RPB
        img_bytes = nb1 * BLOCK_BYTES
        d_img = torch.empty(img_bytes, dtype=torch.uint8, device=dev)
        dbt.check(L.dbt_gen_syn(1234, n1, U1, 0, 0, n1, 0, d_img.data_ptr(), sp))
        h_in, h_out = C.c_void_p(), C.c_void_p()
        dbt.check(L.dbt_host_alloc(C.byref(h_in), img_bytes))
        dbt.check(L.dbt_host_alloc(C.byref(h_out), img_bytes))
        torch.frombuffer((C.c_uint8 * img_bytes).from_address(h_in.value), dtype=torch.uint8).copy_(d_img)
        del d_img
        torch.cuda.empty_cache()
        dbt.check(L.dbt_host_set_chunk_blocks(nb1 // 4))
        nr_, nu_ = C.c_uint64(), C.c_uint64()
        best = None
        for it in range(3):
            t0 = time.perf_counter()
            dbt.check(L.dbt_host_dedup(h_in, nb1, ord(FIELD), h_out, 0, C.byref(nr_), C.byref(nu_)))
            dt = time.perf_counter() - t0
            if it:
                best = dt if best is None else min(best, dt)
        st = (C.c_uint64 * 6)()
        L.dbt_host_ooc_stats(st)
        dbt.check(L.dbt_host_set_chunk_blocks(0))
        assert nu_.value == U1
        out["ooc_dedup_100M_records_4_runs"] = {
            "ms": best * 1e3, "records_per_s": n1 / best, "nunique": nu_.value, "runs": int(st[0]), "output_chunks": int(st[1]),
            "pcie_bytes_approx": img_bytes + 2 * int(st[4]) * BLOCK_BYTES + (U1 // RPB) * BLOCK_BYTES,
            "note": "same workload as the headline, forced out of core: 4 runs deduplicated in-core and parked in pinned host "
                    "memory, one global sort of the resident key columns, 4 output chunks gathered from staged run slices; "
                    "every record crosses PCIe up to four times (in, run out, run in, result out); downloads overlap the next "
                    "chunk's upload on a second stream"}
        L.dbt_host_free(h_in)
        L.dbt_host_free(h_out)
        dbt.check(L.dbt_host_trim())
    except Exception as e:  # noqa: BLE001
        L.dbt_host_set_chunk_blocks(0)
        out["ooc_dedup_100M_records_4_runs"] = {"error": str(e)[:200]}
    # BASELINE configs[3] at ONE GPU: HashJoin field=num, R = 100M x S = 1B records.  S (140 GB) cannot be resident beside
    # R and the output, so it streams through HBM in 10 chunks of 100M records (generated in place between the timed
    # probes, the way a chunk would arrive from storage); R's key column is extracted once and stays resident.
    for kind, label in ((1, "uniform"), (4, "zipf1.1")):
        try:
            nr, ns, D, nchunk = 100_000_000, 1_000_000_000, 100_000_000, 10
            rows_c = ns // nchunk
            nbr, nbc = nr // RPB, rows_c // RPB
            d_r = torch.empty(nbr * BLOCK_BYTES, dtype=torch.uint8, device=dev)
            dbt.check(L.dbt_gen_syn(7, nr, D, 1, 0, nr, 0, d_r.data_ptr(), sp))
            wsb = dbt.dev_ws_bytes(dbt.OP_HASHJOIN, nbr, nbc, FIELD)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            rkeys = torch.empty(nr, dtype=torch.int32, device=dev)
            nrows = C.c_uint64()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            dbt.check(L.dbt_dev_extract_keys_u32(d_r.data_ptr(), nbr, ord(FIELD), rkeys.data_ptr(), ws.data_ptr(), wsb, sp,
                                                 C.byref(nrows)))
            e1.record()
            torch.cuda.synchronize()
            ms_r = e0.elapsed_time(e1)
            del d_r
            torch.cuda.empty_cache()
            d_s = torch.empty(nbc * BLOCK_BYTES, dtype=torch.uint8, device=dev)
            d_o = torch.empty(nbc * BLOCK_BYTES, dtype=torch.uint8, device=dev)
            total, ms_s = 0, 0.0
            k = C.c_uint64()
            for c in range(nchunk + 1):  # chunk 0 runs twice: the first run is the warm-up
                cc = max(c - 1, 0)
                dbt.check(L.dbt_gen_syn(9, ns, D, kind, cc * rows_c, rows_c, 0, d_s.data_ptr(), sp))
                torch.cuda.synchronize()
                e0.record()
                dbt.check(L.dbt_dev_semijoin_keys(rkeys.data_ptr(), nr, d_s.data_ptr(), nbc, ord(FIELD), d_o.data_ptr(), nbc,
                                                  ws.data_ptr(), wsb, sp, C.byref(k)))
                e1.record()
                torch.cuda.synchronize()
                if c:
                    ms_s += e0.elapsed_time(e1)
                    total += k.value
            ms = ms_r + ms_s
            out[f"hashjoin_R100M_S1B_streamed_{label}"] = {
                "probe_tuples_per_s": ns / (ms * 1e-3), "ms": ms, "ms_extract_R_keys": ms_r, "ms_probe_chunks": ms_s,
                "chunks": nchunk, "nres": total, "selectivity": total / ns,
                "note": "BASELINE configs[3] at 1 GPU, device scope: S streams through HBM in 10 chunks of 100M records "
                        "(dbt_dev_semijoin_keys per chunk: bitmap of R's keys, two streaming passes over the chunk), R's keys "
                        "resident; same generator and seeds as the multi-GPU configs[3] leg, so nres must equal that run's "
                        "(632,100,452 uniform / 711,420,513 Zipf(1.1))"}
            del d_s, d_o, ws, rkeys
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            out[f"hashjoin_R100M_S1B_streamed_{label}"] = {"error": str(e)[:200]}
    return out


if __name__ == "__main__":
    sys.exit(main())
