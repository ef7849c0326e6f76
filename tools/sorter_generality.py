#!/usr/bin/env python
"""sorter_generality.py — build time of one block on texts that are NOT i.i.d.: what prefix doubling with closed-form long runs
pays for repeats.  Prints one line per family (device-resident build, second of two runs); nothing here is a bench value.

    python tools/sorter_generality.py [symbols=64000000]
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import gecoz_b200 as G
from gecoz_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64_000_000
rng = np.random.default_rng(5)
ACGT = np.frombuffer(b"ACGT", np.uint8)


def fibonacci(length):
    a, b = b"A", b"AC"
    while len(b) < length:
        a, b = b, b + a
    return np.frombuffer(b[:length], np.uint8).copy()


def families():
    yield "i.i.d. ACGT + 1% N", synth.iid_acgtn(n, 1)
    yield "chr-shaped (N runs of 10 kb .. 7 % of the text)", synth.chromosome_shaped(n, 3)
    yield "tandem repeat, period 10", np.tile(np.frombuffer(b"ACACACACGT", np.uint8), n // 10 + 1)[:n].copy()
    unit = ACGT[rng.integers(0, 4, 1000)]
    yield "tandem repeat, random unit of 1000", np.tile(unit, n // 1000 + 1)[:n].copy()
    half = ACGT[rng.integers(0, 4, n // 2)]
    yield "the same random half twice (one repeat of n/2)", np.concatenate([half, half])
    s = ACGT[rng.integers(0, 4, n)]
    for p in rng.integers(0, n - 70_000, 200):                       # 200 copies of one 50 kb segment: segmental duplications
        s[p:p + 50_000] = s[:50_000]
    yield "200 copies of a 50 kb segment", s
    yield "Fibonacci string over {A, C}", fibonacci(n)
    yield "one symbol (all A)", np.full(n, ord("A"), np.uint8)


for name, seq in families():
    text = synth.block_of([seq])
    d_text = torch.from_numpy(text).cuda()
    shape = G.shape_from_counts(G.symbol_counts(d_text, 0))
    d_gcz = torch.empty(int(shape.size), dtype=torch.uint8, device="cuda")
    d_gcx = torch.empty(G.index_size(len(text), 5), dtype=torch.uint8, device="cuda")
    try:
        for _ in range(2):
            t = G.build_block(0, d_text, len(text), 32, shape, d_gcz, d_gcx)
        print(f"{name:50s} n={len(text):>10d}  total {t['total_ms']:9.2f} ms = {len(seq) / 1e3 / t['total_ms']:8.0f} Mbp/s   first sort {t['sort_initial_ms']:8.2f}  "
              f"refine {t['sort_refine_ms']:9.2f} ({t['refine_rounds']} rounds, {t['unresolved_after_first_sort']} unresolved after the first sort, "
              f"{t['long_runs']} long runs, k={t['symbols_per_key']})", flush=True)
    except G.GczError as ex:
        print(f"{name:50s} refused: {ex}", flush=True)
    del d_text, d_gcz, d_gcx
