"""Synthetic ACGTN workloads of the shapes BASELINE.json names.

All generators are seeded numpy PCG64 streams, upper case only.  There is no network for real genomes, so
bench.py and the tests use these and say so ("data": "synthetic").
"""
from __future__ import annotations

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
N = ord("N")

HG38_LENGTHS = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
                133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
                58617616, 64444167, 46709983, 50818468, 156040895, 57227415, 16569]
HG38_NAMES = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY", "chrM"]


def iid_acgtn(length: int, seed: int, p_n: float = 0.01) -> np.ndarray:
    """cfg1 sequence: i.i.d. A/C/G/T with isolated N of probability p_n."""
    rng = np.random.default_rng(seed)
    seq = ACGT[rng.integers(0, 4, length, dtype=np.uint8)]
    if p_n > 0:
        k = rng.binomial(length, p_n)
        seq[rng.integers(0, length, k)] = N
    return seq


def chromosome_shaped(length: int, seed: int) -> np.ndarray:
    """cfg2/cfg3 sequence: i.i.d. ACGT with hg38-like N runs, scaled to `length` (chr1 = 248 956 422 bp):
    10 kb of N at each end, one 18 Mbp run starting at 40 %, forty 50 kb runs at seeded positions."""
    rng = np.random.default_rng(seed)
    seq = ACGT[rng.integers(0, 4, length, dtype=np.uint8)]
    scale = length / 248956422.0
    tel = max(1, int(10_000 * scale))
    seq[:tel] = N
    seq[length - tel:] = N
    cen = int(18_000_000 * scale)
    start = int(0.4 * length)
    seq[start:start + cen] = N
    gap = max(1, int(50_000 * scale))
    for p in rng.integers(0, max(1, length - gap), 40):
        seq[p:p + gap] = N
    return seq


def block_of(sequences) -> np.ndarray:
    """Generalized string of one block: every sequence followed by '\\0'."""
    out = np.zeros(sum(len(s) + 1 for s in sequences), dtype=np.uint8)
    p = 0
    for s in sequences:
        out[p:p + len(s)] = s
        p += len(s) + 1
    return out


def cfg1_text(length: int = 16_000_000) -> np.ndarray:
    return block_of([iid_acgtn(length, seed=1)])


def cfg2_text(length: int = 248_956_422, seed: int = 3) -> np.ndarray:
    return block_of([chromosome_shaped(length, seed)])


def hg38_shaped_records(scale: float = 1.0):
    """cfg3: 25 sequences with the hg38 lengths (optionally scaled down for tests), seed 4 + i."""
    recs = []
    for i, (name, length) in enumerate(zip(HG38_NAMES, HG38_LENGTHS)):
        ln = max(8, int(length * scale))
        recs.append((name, chromosome_shaped(ln, 4 + i)))
    return recs


def patterns(text: np.ndarray, count: int, min_len: int, max_len: int, seed: int):
    """Half sampled from the text (windows containing N or a separator are re-drawn), half i.i.d. ACGT.
    Returns (bytes uint8[total], offsets int64[count + 1])."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(min_len, max_len + 1, count).astype(np.int64)
    off = np.zeros(count + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    data = ACGT[rng.integers(0, 4, int(off[-1]), dtype=np.uint8)]
    n = len(text)
    from_text = np.flatnonzero(rng.random(count) < 0.5)
    bad_pos = np.flatnonzero((text == N) | (text == 0))  # sorted; a window [s, e) is clean when no position falls in it

    def touches_bad(s, e):
        return np.searchsorted(bad_pos, s) != np.searchsorted(bad_pos, e)

    starts = rng.integers(0, max(1, n - max_len - 1), len(from_text))
    tlen = lens[from_text]
    check = np.arange(len(starts))                       # windows not yet known to be clean
    for _ in range(8):                                   # re-draw windows that touch N / '\0'
        s_c = starts[check]
        check = check[touches_bad(s_c, np.minimum(s_c + tlen[check], n))]
        if not len(check):
            break
        starts[check] = rng.integers(0, max(1, n - max_len - 1), len(check))
    clean = np.ones(len(starts), dtype=bool)
    if len(check):
        s_c = starts[check]
        clean[check] = ~touches_bad(s_c, np.minimum(s_c + tlen[check], n))
    sel, st = from_text[clean], starts[clean]
    for c in range(0, len(sel), 1 << 18):                # bounded temporaries: 256 K patterns at a time
        s_c, st_c = sel[c:c + (1 << 18)], st[c:c + (1 << 18)]
        ln = lens[s_c]
        first = np.zeros(len(s_c), dtype=np.int64)
        np.cumsum(ln[:-1], out=first[1:])
        within = np.arange(int(ln.sum()), dtype=np.int64) - np.repeat(first, ln)
        data[np.repeat(off[s_c], ln) + within] = text[np.repeat(st_c, ln) + within]
    return data, off
