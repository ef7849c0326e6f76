"""Tuning aid: device time of the onesweep digit passes over (u64, u32) pairs with random keys:  python tools/sortbench.py [n] [bits]"""
import ctypes as C
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import gecoz_b200 as G
from gecoz_b200 import _native as N

n = int(sys.argv[1]) if len(sys.argv) > 1 else 248_956_423
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 63
torch.manual_seed(0)
keys0 = torch.randint(0, 2 ** 62, (n,), dtype=torch.int64, device="cuda")
if bits < 62:
    keys0 &= (1 << bits) - 1
vals0 = torch.arange(n, dtype=torch.int32, device="cuda")
res = []
for it in range(3):
    keys, vals = keys0.clone(), vals0.clone()
    N.check(G.lib().gcz_dbg_sort_pairs(0, N.ptr(keys), N.ptr(vals), n, 0, bits))
    t = N.BuildTiming()
    G.lib().gcz_last_build_timing(C.byref(t))
    res.append((t.radix_ms, t.radix_launches))
ok = bool((keys[1:] >= keys[:-1]).all())
ms, passes = res[-1]
print(f"n={n} bits={bits} passes={passes} ms={ms:.3f} "
      f"per_pass={ms / passes:.3f} GB/s={24 * n * passes / ms / 1e6:.0f} sorted={ok}")
