"""Summarise `ncu -i report.ncu-rep --page raw --csv` into the JSON list kept as profiles/r01_ncu_full_*.json.

usage: ncu -i gpurun_out/x.ncu-rep --page raw --csv > /tmp/raw.csv; python profiles/summarize_ncu.py /tmp/raw.csv x.ncu-rep > out.json
"""
import csv
import json
import re
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]


def main():
    raw, rep = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    with open(raw, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.reader(lines))
    header, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(header, r))
        name = re.sub(r"\(.*", "", d.get("Kernel Name", "")).replace("void ", "").split("::")[-1]
        e = {"kernel": name, "report": rep}
        for k in KEEP:
            if k in d:
                e[k] = f"{d[k]} {units[header.index(k)]}".strip()
        out.append(e)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
