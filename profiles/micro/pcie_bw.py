#!/usr/bin/env python
"""Host<->device copy ceiling of the box: pinned cudaMemcpyAsync H2D, D2H and both at once on 1, 2, 4, ... all visible
GPUs concurrently (one process, one stream per direction per GPU).  Prints one JSON line per configuration.
VERDICT r1 item 3: what aggregate PCIe rate can N ranks get at best, to judge the e2e numbers against."""
import json
import sys
import time

import torch


def run(ngpu: int, gb: float, mode: str, reps: int = 3):
    n = int(gb * (1 << 30))
    bufs = []
    for d in range(ngpu):
        torch.cuda.set_device(d)
        h_in = torch.empty(n, dtype=torch.uint8).pin_memory() if mode in ("h2d", "both") else None
        h_out = torch.empty(n, dtype=torch.uint8).pin_memory() if mode in ("d2h", "both") else None
        d_a = torch.empty(n, dtype=torch.uint8, device=f"cuda:{d}")
        d_b = torch.empty(n, dtype=torch.uint8, device=f"cuda:{d}")
        bufs.append((h_in, h_out, d_a, d_b, torch.cuda.Stream(device=d), torch.cuda.Stream(device=d)))
    best = 1e9
    for _ in range(reps + 1):
        for d in range(ngpu):
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for d, (h_in, h_out, d_a, d_b, s1, s2) in enumerate(bufs):
            if h_in is not None:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if h_out is not None:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
        for d in range(ngpu):
            torch.cuda.synchronize(d)
        best = min(best, time.perf_counter() - t0)
    per_dir = ngpu * n / best / 1e9
    return {"gpus": ngpu, "mode": mode, "gb_per_gpu_per_direction": gb, "seconds": round(best, 4),
            "aggregate_gbs_per_direction": round(per_dir, 1), "aggregate_gbs_total": round(per_dir * (2 if mode == "both" else 1), 1)}


if __name__ == "__main__":
    gb = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
    nmax = torch.cuda.device_count()
    k = 1
    while k <= nmax:
        for mode in ("h2d", "d2h", "both"):
            print(json.dumps(run(k, gb, mode)), flush=True)
        k *= 2
