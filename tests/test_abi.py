"""CPU tests of the boundary: the shared library loads, exports every symbol the headers declare, the
pure-host counter arithmetic is right, and there is no CPU fallback."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_everything_the_c_header_declares(dbt):
    hdr = open(os.path.join(ROOT, "include", "dbt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dbt_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    L = dbt.lib()
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    assert declared == set(dbt.C_ABI_SYMBOLS), declared ^ set(dbt.C_ABI_SYMBOLS)


def test_drop_in_entry_points_have_the_reference_mangled_names(dbt):
    # the symbols main.o imports from the reference's DatabaseProject.o
    L = dbt.lib()
    for name, sym in dbt.CXX_ENTRY_POINTS.items():
        assert hasattr(L, sym), (name, sym)
    out = subprocess.run(["nm", "-D", "--defined-only", dbt.LIB_PATH], capture_output=True, text=True).stdout
    for sym in dbt.CXX_ENTRY_POINTS.values():
        assert re.search(rf" T {sym}$", out, flags=re.M), sym


def test_contract_header_layout_compiles_and_matches(tmp_path):
    src = tmp_path / "layout.cpp"
    src.write_text(
        '#include "dbtproj.h"\n#include <cstddef>\n#include <cstdio>\n'
        'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(record_t), sizeof(block_t), offsetof(record_t,num),'
        ' offsetof(record_t,str), offsetof(record_t,valid), offsetof(record_t,dummy1), offsetof(block_t,entries),'
        ' offsetof(block_t,dummy)); return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["g++", "-std=c++11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    assert subprocess.check_output([str(exe)], text=True).split() == ["140", "14016", "4", "8", "128", "132", "8", "14012"]


def test_counters_are_pure_host_arithmetic(dbt, orc):
    for B, M in [(10000, 64), (10000, 101), (200, 8), (600, 8), (600, 3), (300, 16), (1, 3), (1000000, 64)]:
        assert dbt.sort_counters(B, M) == orc.sort_counters(B, M)
    L = dbt.lib()
    assert L.dbt_dedup_nios(300, 64, 8698) == orc.dedup_nios(300, 64, 8698) == 707
    assert L.dbt_hashjoin_nios(300, 300, 64, 28972) == orc.hashjoin_nios(300, 300, 64, 28972)
    res = (C.c_uint64 * 4)(8386, 8698, 8682, 173)
    info = {"nres": 8386, "nunique_R": 8698, "nunique_S": 8682, "later_reads": 173}
    assert L.dbt_mergejoin_nios(300, 300, 64, res) == orc.mergejoin_nios(300, 300, 64, info)
    with pytest.raises(dbt.DbtError):
        dbt.sort_counters(100, 2)  # "The buffer size is too small!"


def test_no_cpu_fallback_without_a_device(dbt):
    L = dbt.lib()
    if L.dbt_device_count() > 0:
        pytest.skip("a CUDA device is visible here")
    import numpy as np

    buf = np.zeros(14016, np.uint8)
    out = np.zeros(14016, np.uint8)
    n = C.c_uint64()
    rc = L.dbt_host_mergesort(buf.ctypes.data, 1, ord("1"), out.ctypes.data, 0, C.byref(n))
    assert rc == -2  # DBT_ERR_CUDA, loudly
    assert b"no CPU fallback" in L.dbt_last_error() or b"CUDA" in L.dbt_last_error()


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "database-technology-algorithms_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                # code references only (a comment may cite the oracle file that mirrors the generator)
                if re.search(r"#include[^\n]*oracle|^\s*(from|import)\s+oracle|pyoracle|liboracle|ref_runner|dlopen", text, flags=re.M):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders


def test_integration_notes_only_name_symbols_the_header_declares():
    """INTEGRATION.md is what a maintainer binds against: every dbt_* function it names is declared in include/dbt_b200.h."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    hdr = open(os.path.join(ROOT, "include", "dbt_b200.h")).read()
    declared = set(re.findall(r"\bdbt_[a-z0-9_]+\b", hdr))
    named = {t for t in re.findall(r"\bdbt_[a-z0-9_]+\b", doc) if not t.endswith("_")}  # (dbt_dist_* etc. are families, not names)
    named -= {"dbt_main", "dbt_b200", "dbt_oracle", "dbt_status"}                         # a binary, the header, a file, an enum
    assert not (named - declared), sorted(named - declared)


def test_bench_reference_arm_prints_a_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-rows", "200000"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    import json

    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    if "unavailable" not in line:
        assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] == 1
        assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0


def test_reference_main_links_against_the_library_unchanged(dbt, tmp_path):
    """The drop-in claim at link level: the reference's own main.cpp (untouched, compiled from where it
    lies) resolves MergeJoin/HashJoin from libdbt_b200.so.  Only where /root/reference exists."""
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "main.cpp")):
        pytest.skip("no /root/reference here")
    exe = tmp_path / "dbt_ref_main"
    libdir = os.path.dirname(dbt.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++11", "-w", "-I", ref, os.path.join(ref, "main.cpp"), "-L", libdir, "-ldbt_b200",
                           f"-Wl,-rpath,{libdir}", "-o", str(exe)])
    und = subprocess.check_output(["nm", "-u", str(exe)], text=True)
    assert dbt.CXX_ENTRY_POINTS["MergeJoin"] in und and dbt.CXX_ENTRY_POINTS["HashJoin"] in und
