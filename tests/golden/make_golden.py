#!/usr/bin/env python
"""Generates tests/golden/ref_golden.{npz,json} by RUNNING THE UNMODIFIED REFERENCE
(oracle/_ref/ref_runner, built from /root/reference by oracle/Makefile) on seeded inputs.

The reference has no tests or golden vectors of its own, so these fixtures are
the pinned record of what the reference binary does on this toolchain (g++ 13.3, glibc 2.39):
they travel to the GPU box, where /root/reference does not exist.

    python tests/golden/make_golden.py        # needs /root/reference (or a prebuilt oracle/_ref/ref_runner)

What is frozen (inputs are regenerated from the seed by oracle/dbt_oracle.c orc_gen_ref = main.cpp:41-77):
  * sort_f<field>: recids of REF's MergeSort output, ties canonicalised by recid (REFc),
    for a single-merge-phase configuration (npasses = 2: REF is lossless there);
  * sort3p_f1: the same for npasses = 3, where REF loses a few rows (D1/D2): recids REF did output;
  * hjoin_f<field>: recids REF's HashJoin emitted, in its output order (S file order);
  * hjoinm_f3: HashJoin field '3' with REAL multiplicities: a small num modulus makes the "Hola" rows of R (row 1 of
    every block, main.cpp:57-61) share (num, str) keys, so S's "Hola" rows are emitted once per matching R row
    (DatabaseProject.cpp:616-629).  The seed is searched so that no S row's group of emissions crosses an output-block
    boundary: only then does REF stay inside its output block (defect D10) and its result is usable;
  * counters: nsorted_segs / npasses / nios / nunique / nres as printed by REF.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SEED, NBLOCKS = 42, 60          # 6,000 rows per file
SEED3, NBLOCKS3, NMEM3 = 7, 200, 8  # 20,000 rows, M=8 => 25 -> 4 -> 1 runs: npasses 3


def main():
    orc.build()
    assert orc.ref_available(), "build oracle/_ref/ref_runner first (make -C oracle)"
    arrays, meta = {}, {"seed": SEED, "nblocks": NBLOCKS, "seed3": SEED3, "nblocks3": NBLOCKS3, "nmem3": NMEM3,
                        "toolchain": "g++ 13.3 / glibc 2.39", "counters": {}}
    f1, f2 = orc.gen_ref(SEED, NBLOCKS)
    arrays["input_f1_num"] = orc.rows_of(f1)["num"].copy()
    arrays["input_f2_num"] = orc.rows_of(f2)["num"].copy()
    for field in "0123":
        info, out, _ = orc.run_ref("sort", field, 64, f1)
        refc = orc.canonicalise_ties(out[out["nreserved"] > 0], field)
        arrays[f"sort_f{field}"] = orc.rows_of(refc)["recid"].copy()
        meta["counters"][f"sort_f{field}"] = {"nsorted_segs": info["a"], "npasses": info["b"], "nios": info["nios"],
                                              "outfile": info["outfile"], "nmem": 64}
        info, out, _ = orc.run_ref("dedup", field, 64, f1)
        meta["counters"][f"dedup_f{field}"] = {"nunique": info["a"], "nios": info["nios"], "nmem": 64}
        info, out, _ = orc.run_ref("hjoin", field, 64, f1, f2)
        arrays[f"hjoin_f{field}"] = orc.rows_of(out, info["a"])["recid"].copy()
        meta["counters"][f"hjoin_f{field}"] = {"nres": info["a"], "nios": info["nios"], "nmem": 64}
    # field '3' with multiplicities, on a seed where REF does not overflow its output block (D10)
    MULT_NB, MULT_MOD = 60, 12
    for seed in range(1, 2000):
        m1, m2 = orc.gen_ref(seed, MULT_NB, num_mod=MULT_MOD)
        ids = orc.rows_of(orc.hashjoin(m1, m2, "3"))["recid"]
        if len(ids) < 150:
            continue
        cuts = np.flatnonzero(np.diff(ids)) + 1                      # group starts (one group per emitting S row)
        starts = np.concatenate([[0], cuts]); ends = np.concatenate([cuts, [len(ids)]])
        mult = ends - starts
        straddle = ((starts // 100) != ((ends - 1) // 100)).any()
        if mult.max() >= 3 and not straddle:
            break
    else:
        raise SystemExit("no usable seed for the multiplicity fixture")
    info, out, _ = orc.run_ref("hjoin", "3", 64, m1, m2)
    assert info.get("exit", 1) == 0 and info["a"] == len(ids), (info, len(ids))
    arrays["hjoinm_f3"] = orc.rows_of(out, info["a"])["recid"].copy()
    meta["mult"] = {"seed": seed, "nblocks": MULT_NB, "num_mod": MULT_MOD, "nres": info["a"], "nios": info["nios"], "nmem": 64,
                    "max_multiplicity": int(mult.max()), "emitting_s_rows": int(len(mult))}
    g1 = orc.gen_ref(SEED3, NBLOCKS3, two=False)
    info, out, _ = orc.run_ref("sort", "1", NMEM3, g1)
    refc = orc.canonicalise_ties(out[out["nreserved"] > 0], "1")
    arrays["sort3p_f1"] = orc.rows_of(refc)["recid"].copy()
    meta["counters"]["sort3p_f1"] = {"nsorted_segs": info["a"], "npasses": info["b"], "nios": info["nios"],
                                     "rows_out": int(len(arrays["sort3p_f1"])), "nmem": NMEM3}
    # counter table straight from REF, small sizes only (field '0': sorted input, fast)
    table = []
    for nb, M in [(200, 8), (300, 16), (600, 8), (600, 3), (100, 101), (64, 64), (65, 64), (1, 3)]:
        g = orc.gen_ref(1, nb, two=False)
        info, out, _ = orc.run_ref("sort", "0", M, g)
        table.append({"nblocks": nb, "nmem": M, "nsorted_segs": info["a"], "npasses": info["b"], "nios": info["nios"],
                      "rows_out": int(orc.count_rows(out))})
    meta["counter_table"] = table
    np.savez_compressed(os.path.join(HERE, "ref_golden.npz"), **arrays)
    with open(os.path.join(HERE, "ref_golden.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", {k: v.shape for k, v in arrays.items()})
    print(json.dumps(meta["counters"], indent=1)[:1500])
    print(table)


if __name__ == "__main__":
    main()
