#!/usr/bin/env python
"""build_once.py — N device-resident builds of the chr1-shaped block and nothing else: the command ncu wraps.

    python tools/build_once.py [N=2] [length=248956422]
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import gecoz_b200 as G
from gecoz_b200 import synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
length = int(sys.argv[2]) if len(sys.argv) > 2 else 248_956_422
text = synth.cfg2_text(length, seed=3)
n = len(text)
d_text = torch.from_numpy(text).cuda()
shape = G.shape_from_counts(G.symbol_counts(d_text, 0))
d_gcz = torch.empty(int(shape.size), dtype=torch.uint8, device="cuda")
d_gcx = torch.empty(G.index_size(n, 5), dtype=torch.uint8, device="cuda")
for _ in range(reps):
    t = G.build_block(0, d_text, n, 32, shape, d_gcz, d_gcx)
print({k: round(float(t[k]), 3) for k in ("total_ms", "sort_initial_ms", "sort_refine_ms", "bwt_hswt_ms", "ssa_ms", "kernel_launches")})
