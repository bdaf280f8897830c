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
N>1 (torchrun): every rank holds its own 100M-row shard of one global relation; sample-sort splitters,
one all-to-all of (key, row) pairs and one of the surviving records (see dist_ops.py).
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
        "workload": f"EliminateDuplicates field=num, {args.rows} records per GPU, 10% duplicate rows (BASELINE configs[1]: 100M records, 1 B200)",
        "rows_per_gpu": args.rows, "distinct_keys_per_gpu": (args.rows * 9) // 10, "nmem_blocks": NMEM,
        "key_bits": 32, "record_bytes": 140, "cache": "input image (14 GB at 100M rows) is far larger than the 126 MB L2; no flush needed",
        "parallelism": "1 GPU" if world == 1 else f"{world} GPUs, sample-sort range partition + all-to-all",
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
    print(json.dumps(line))
    return 0


def bench_dist(args, dbt, rank, world, local_rank):
    """N>1: one global relation of N x rows records (10% duplicate rows globally), rank r holds rows
    [r*rows, (r+1)*rows).  A step = distributed EliminateDuplicates: range-partition by sample-sort
    splitters, one all-to-all of record images over NVLink, local dedup; the concatenation of the ranks'
    outputs is the globally sorted duplicate-free file (weak scaling: rows per GPU fixed)."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    dmod = importlib.import_module("database-technology-algorithms_b200.dist")
    L = dbt.lib()
    dev = torch.device("cuda", local_rank)
    ops = dmod.LocalOps(dev)
    d = dmod.DistOps(ops)
    peak, peak_src = load_peaks()
    n = args.rows
    n_total = n * world
    U = (n_total * 9) // 10
    nblocks = (n + RPB - 1) // RPB
    img_bytes = nblocks * BLOCK_BYTES
    stream = torch.cuda.current_stream()
    d_in = torch.empty(img_bytes, dtype=torch.uint8, device=dev)
    dbt.check(L.dbt_gen_syn(42, n_total, U, 0, rank * n, n, 0, d_in.data_ptr(), stream.cuda_stream))
    torch.cuda.synchronize()

    def step():
        out, info = d.dedup(d_in, nblocks, FIELD)
        return out, info

    for _ in range(max(args.warmup, 3)):
        out, info = step()
    got = d.total(info["out_rows"])
    assert got == U, f"wrong global result: {got} unique rows, expected {U}"
    del out

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        L.dbt_stage_timing_enable(1)
        L.dbt_stage_timing_reset()
    launches0 = L.dbt_kernel_launches()
    xms, xbytes = [], 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        out, info = step()
        ev = d.last_exchange.get("gather_events") or d.last_exchange["events"]
        xbytes = d.last_exchange.get("remote_record_bytes_read") or d.last_exchange["bytes_sent_remote"]
        xms.append(ev)
        del out
    e1.record(stream)
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / args.steps
    launches = int(L.dbt_kernel_launches() - launches0)
    a2a_ms = sum(a.elapsed_time(b) for a, b in xms) / len(xms)
    a2a = torch.tensor([a2a_ms, float(xbytes)], device=dev, dtype=torch.float64)
    dist.all_reduce(a2a, op=dist.ReduceOp.MAX)
    stages, clocks = {}, None
    timeline = None
    if rank == 0 and "timeline" in d.last_exchange:
        tl = d.last_exchange["timeline"]
        timeline = {name: round(tl[0][1].elapsed_time(e), 3) for name, e in tl}
        pe = d.last_exchange["events"]
        timeline["push_start"] = round(tl[0][1].elapsed_time(pe[0]), 3)
        timeline["push_end"] = round(tl[0][1].elapsed_time(pe[1]), 3)
    if rank == 0:
        stages = dbt.stage_report()
        L.dbt_stage_timing_enable(0)
        clocks = sampler.stop()
    value = n_total / (ms_step * 1e-3)

    # e2e: pinned host shard in, pinned host result out, copies inside the timed region (per rank)
    e2e = None
    if not args.no_e2e:
        try:
            import psutil

            avail = psutil.virtual_memory().available
        except Exception:  # noqa: BLE001
            avail = 0
        need = 2 * img_bytes * world * 1.25
        if avail > need:
            h_in = torch.empty(img_bytes, dtype=torch.uint8).pin_memory()
            h_out = torch.empty(img_bytes + img_bytes // 4, dtype=torch.uint8).pin_memory()
            h_in.copy_(d_in)
            d_stage = torch.empty_like(d_in)

            def e2e_step():
                d_stage.copy_(h_in, non_blocking=True)
                o, inf = d.dedup(d_stage, nblocks, FIELD)
                nb_out = (inf["out_rows"] + RPB - 1) // RPB
                h_out[: nb_out * BLOCK_BYTES].copy_(o[: nb_out * BLOCK_BYTES], non_blocking=True)
                torch.cuda.synchronize()
                return nb_out * BLOCK_BYTES

            e2e_step()
            dist.barrier()
            t0 = time.perf_counter()
            ob = 0
            for _ in range(args.steps):
                ob = e2e_step()
            dist.barrier()
            sec = torch.tensor([(time.perf_counter() - t0) / args.steps], device=dev, dtype=torch.float64)
            dist.all_reduce(sec, op=dist.ReduceOp.MAX)
            e2e = {"value": n_total / float(sec.item()), "unit": "records/s", "h2d_bytes_per_step": img_bytes * world,
                   "d2h_bytes_per_step": ob * world, "ms_per_step": float(sec.item()) * 1e3,
                   "api": "pinned host shard -> DistOps.dedup (C-ABI kernels + NCCL all-to-all) -> pinned host result, per rank"}
        else:
            e2e = {"value": None, "unit": "records/s", "h2d_bytes_per_step": img_bytes * world, "d2h_bytes_per_step": None,
                   "skipped": f"needs {need/1e9:.0f} GB of pinned host memory, {avail/1e9:.0f} GB available"}

    if rank == 0:
        dom = max(stages.items(), key=lambda kv: kv[1][0])[0] if stages else None
        roofline = None
        if dom == "record_gather":
            # per step on this rank: the P per-destination gathers (n rows) + the final gather (rows received, ~0.9 n)
            ms_dom = stages[dom][0] / args.steps
            rows_moved = n + info["out_rows"]
            achieved = 280.0 * rows_moved / (ms_dom * 1e-3) / 1e9
            roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": load_traffic(dom), "peak_source": peak_src, "ms_per_step_all_launches": ms_dom,
                        "algorithmic_bytes_per_step": 280.0 * rows_moved}
        nvlink = {"exchange": d.last_exchange.get("mode"),
                  "what": "record bytes crossing NVLink per GPU per step, and the duration of the phase that moves them (max over ranks)",
                  "bytes_sent_per_gpu": float(a2a[1].item()), "all_to_all_ms": float(a2a[0].item()),
                  "achieved_gbs_per_direction": float(a2a[1].item()) / (float(a2a[0].item()) * 1e-3) / 1e9 if a2a[0].item() > 0 else None,
                  "peak_gbs_per_direction": 770.0, "peak_source": "B200_PROFILING.md measured peer copy"}
        line = {
            "metric": "records_per_second", "value": value, "unit": "records/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": workload_config(args, world),
            "e2e": e2e, "roofline": roofline, "nvlink": nvlink, "cpu_baseline": None, "clocks": clocks,
            "gpu_launches": launches, "stage_ms_per_step": {k: round(v[0] / args.steps, 4) for k, v in stages.items()},
            "timeline_ms_last_step_rank0": timeline,
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
    return 0


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
    for kind, label in ((1, "uniform"), (2, "skewed")):
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
            probe_ms = rep["hash_probe"][0]
            sel = k / ns
            out[f"hashjoin_R100M_S400M_{label}"] = {
                "probe_tuples_per_s": ns / (ms * 1e-3), "ms": ms, "nres": k, "selectivity": sel,
                "stage_ms": {a: round(b[0], 3) for a, b in rep.items()},
                "probe_kernel_hbm_frac_of_measured": (8 + 4 * sel) * ns / (probe_ms * 1e-3) / 1e9 / peak,
                "note": "semi-join (reference semantics): S rows in S order whose key is in keys(R); probe kernel = "
                        "L2-resident bitmap test because keys(R) span < 2^29; whole operator includes extraction of "
                        "both images and the gather of the matching S records"}
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
    for kind, label in ((1, "uniform"), (2, "skewed")):
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
                        "(dbt_dev_semijoin_keys per chunk: extraction, bitmap build + probe, compaction, gather of the "
                        "matching S records), R's keys resident; same generator and seeds as the 8-GPU run of "
                        "profiles/dist_configs.py, so nres must equal that run's"}
            del d_s, d_o, ws, rkeys
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            out[f"hashjoin_R100M_S1B_streamed_{label}"] = {"error": str(e)[:200]}
    return out


if __name__ == "__main__":
    sys.exit(main())
