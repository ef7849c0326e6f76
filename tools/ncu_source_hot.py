#!/usr/bin/env python
"""Hot source lines of a kernel from an ncu report: instructions executed and stall samples per source line.

    ncu -i rep.ncu-rep --page source --csv --print-source cuda ... > src.csv   (or the default SASS page with -lineinfo)
    python tools/ncu_source_hot.py src.csv [top]
Works on the SASS listing: groups by opcode when no source column mapping exists."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ci, si, ii = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot_i = tot_s = 0
items = []
for r in rows[2:]:
    if len(r) <= ii:
        continue
    try:
        ins, smp = float(r[ii] or 0), float(r[si] or 0)
    except ValueError:
        continue
    tot_i += ins
    tot_s += smp
    items.append((ins, smp, r[0], r[ci]))
print(f"total instructions {tot_i:.3e}, samples {tot_s:.0f}")
byop = collections.Counter()
for ins, smp, _, src in items:
    byop[src.split()[0] if not src.startswith("@") else src.split()[1]] += ins
print("by opcode:", ", ".join(f"{k} {100 * v / tot_i:.1f}%" for k, v in byop.most_common(14)))
print("-- top by stall samples")
for ins, smp, addr, src in sorted(items, key=lambda x: -x[1])[:top]:
    print(f"{100 * smp / max(tot_s, 1):5.1f}% smp {100 * ins / max(tot_i, 1):5.2f}% ins  {addr[-5:]}  {src[:110]}")
