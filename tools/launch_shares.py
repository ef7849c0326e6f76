#!/usr/bin/env python
"""Per-kernel shares of ONE build step from an ncu launch list (--metrics gpu__time_duration.sum --csv).

    python tools/launch_shares.py gpurun_out/launches.csv [step_index]
A step runs from one symbol_histogram_kernel (head of gcz_build_block) to the next."""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for x in csv.DictReader(lines):
    rows.append((x["Kernel Name"], float(x["Metric Value"].replace(",", ""))))
heads = [i for i, x in enumerate(rows) if "symbol_histogram" in x[0]]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
st = rows[heads[which]:heads[which + 1]]
agg = collections.OrderedDict()
for name, v in st:
    n = re.sub(r"^.*::", "", re.sub(r"\(.*", "", name))
    agg.setdefault(n, [0, 0.0])
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches | ms | share |\n|---|---|---|---|")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {n} | {c} | {v * 1e-6:.3f} | {100 * v / tot:.1f} % |")
print(f"| total | {len(st)} | {tot * 1e-6:.3f} | |")
