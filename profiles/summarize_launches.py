"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) into per-kernel totals and shares.

usage: python profiles/summarize_launches.py raw.csv "<command that was profiled>" > summary.csv
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    raw, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = []
    with open(raw, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    header = None
    for r in rd:
        if header is None:
            if "Kernel Name" in r:
                header = r
            continue
        rows.append(dict(zip(header, r)))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"<.*", "", r["Kernel Name"]).split("(")[0].split("::")[-1].replace("void ", "").strip()
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    op_total = sum(v[1] for k, v in agg.items() if k != "gen_syn_kernel")
    print(f"# ncu launch list; command: {cmd}")
    print("# cold-cache serialised times: compare SHARES with bench.py's live stage timing, not absolutes")
    print("kernel,launches,total_ms,avg_ms,share_of_operator")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        share = 0.0 if k == "gen_syn_kernel" else ms / op_total
        print(f"{k},{n},{ms:.3f},{ms / n:.3f},{share:.3f}")


if __name__ == "__main__":
    main()
