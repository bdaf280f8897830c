"""Builds libdbt_b200.so in-tree for sm_100a (and nothing else) with nvcc.

Usage: python build.py [--force]   (also called by __graft_entry__.build()).
The .so lands next to this file so it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libdbt_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr"]
CFLAGS += os.environ.get("DBT_NVCC_EXTRA", "").split()  # tuning hook, e.g. -DDBT_LOOKBACK_WINDOW=8


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")) and not f.startswith("dbt_main"))


def _newer(a: str, deps) -> bool:
    if not os.path.exists(a):
        return False
    t = os.path.getmtime(a)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers += [os.path.join(HERE, "..", "include", f) for f in ("dbt_b200.h", "dbtproj.h")]
    jobs = []
    objs = []
    for f in sources():
        src = os.path.join(CSRC, f)
        obj = os.path.join(OBJ, os.path.splitext(f)[0] + ".o")
        objs.append(obj)
        if force or not _newer(obj, [src] + headers):
            cmd = [NVCC] + ARCH + CFLAGS + ["-x", "cu", "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), p.stdout, p.stderr))
        return p.stderr

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for out in ex.map(run, jobs):
            if verbose and out:
                print(out)
    if jobs or not os.path.exists(LIB):
        run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"])
    # the main.cpp-compatible driver (links against the .so)
    drv_src = os.path.join(CSRC, "dbt_main.cpp")
    drv = os.path.join(HERE, "dbt_main")
    if os.path.exists(drv_src) and (force or not _newer(drv, [drv_src, LIB])):
        run(["g++", "-O2", "-std=c++17", "-I", os.path.join(HERE, "..", "include"), drv_src, "-o", drv,
             "-L", HERE, "-ldbt_b200", "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
