#!/usr/bin/env python
"""Key numbers of an ncu report, one block per profiled launch.

    ncu -i rep.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv
or  python tools/ncu_summary.py rep.ncu-rep          (runs ncu -i itself)

Duration, grid / block / registers, DRAM bytes, instructions, issue-slot and warp occupancy, L2 hit rate, shared-memory
bank conflicts, and the stall reasons per issued instruction above 0.3 — the metrics the tuning log of this repo quotes.
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "  of them bank conflicts"),
]


def main() -> None:
    src = sys.argv[1]
    if src.endswith(".ncu-rep"):
        text = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    else:
        text = open(src).read()
    rows = list(csv.reader(io.StringIO("".join(l + "\n" for l in text.splitlines() if not l.startswith("==")))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for d in data:
        print(d[col["Kernel Name"]][:110])
        for key, label in WANT:
            if key in col:
                print(f"  {label:28s} {d[col[key]]:>18s} {units[col[key]]}")
        stalls = []
        for h, i in col.items():
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(d[i].replace(",", ""))
                except ValueError:
                    continue
                if v > 0.3:
                    stalls.append((v, h.split("issue_stalled_")[-1].replace("_per_issue_active.ratio", "")))
        print("  stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)))


if __name__ == "__main__":
    main()
