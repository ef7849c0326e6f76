#!/usr/bin/env python
"""query_once.py — a few count / find launches on a synthetic multi-block index, for ncu captures of the query kernels.

    python tools/query_once.py small     one 40 Mbp block: the rank sectors (~25 MB) stay in the 126 MB L2
    python tools/query_once.py large     six 200 Mbp blocks: ~0.75 GB of rank sectors, far beyond L2
Prints the device time of the count launches (gcz_last_query_stats) and the wall clock of the find call.
"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import gecoz_b200 as G
from gecoz_b200 import synth

mode = sys.argv[1] if len(sys.argv) > 1 else "small"
sizes = [40_000_000] if mode == "small" else [200_000_000] * 6
gssas, texts = [], []
for i, ln in enumerate(sizes):
    text = synth.block_of([synth.chromosome_shaped(ln, 90 + i)])
    shape = G.shape_from_counts(G.symbol_counts(text))
    gcz = torch.empty(int(shape.size), dtype=torch.uint8, device="cuda")
    gcx = torch.empty(G.index_size(len(text), 5), dtype=torch.uint8, device="cuda")
    G.build_block(0, text, len(text), 32, shape, gcz, gcx)
    gssas.append(G.GSSA.open(0, gcz, len(text), gcx))
    texts.append(text)
data, off = synth.patterns(texts[0], 2_000_000, 15, 100, seed=5)
d_data, d_off = torch.from_numpy(data).cuda(), torch.from_numpy(off).cuda()
tot = torch.zeros(len(off) - 1, dtype=torch.int64, device="cuda")
for _ in range(3):
    G.count_totals(gssas, d_data, d_off, tot)
st = G.last_query_stats()
cs = G.count_stats(gssas, d_data, d_off)
k = 500_000
t0 = time.perf_counter()
block_off, n_hits = G.find_multi(gssas, d_data[:int(off[k])], d_off[:k + 1], copy=False)
sec = time.perf_counter() - t0
print(f"{mode}: {len(sizes)} block(s), index {cs['index_bytes'] / 1e6:.0f} MB; count 2M patterns x {len(sizes)} blocks {st['kernel_ms']:.3f} ms "
      f"({cs['rank_sectors'] * 32 / st['kernel_ms'] / 1e6:.0f} GB/s of rank sectors, {cs['rank_sectors']} sectors, {cs['reference_rank_calls']} reference rank calls); "
      f"find 500k patterns: {n_hits} hits in {sec * 1e3:.1f} ms")
