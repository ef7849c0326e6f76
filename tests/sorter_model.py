"""A numpy model of the GPU suffix sorter's ALGORITHM (gecoz_b200/csrc/suffix_sort.cu), step for step: base-(sigma+1)
keys of k symbols, one stable sort, groups, the closed-form order of suffixes inside long runs of one symbol (side the
run ends on, run length left), prefix doubling with sorted-group elimination at offsets h + (r - k), ranks of suffixes
final after the first sort taken from their slot.  It exists to fuzz the algorithm's invariants on the CPU with forced
small k (thousands of adversarial texts); the CUDA code itself is checked against the oracle in the GPU tests."""
from __future__ import annotations

import numpy as np


def suffix_array_model(text: np.ndarray, k: int | None = None, use_runs: bool = True) -> np.ndarray:
    n = len(text)
    present = np.unique(text)
    code_of = np.zeros(256, np.int64)
    code_of[present] = np.arange(1, len(present) + 1)
    sigma = len(present)
    radix = sigma + 1
    if k is None:
        k = max(1, int(np.floor(62 / np.log2(radix))))
    codes = np.concatenate([code_of[text], np.zeros(k, np.int64)])          # 0 = past the end
    keys = np.zeros(n, dtype=object)
    for j in range(k):
        keys = keys * radix + codes[j:j + n]
    order = np.array(sorted(range(n), key=lambda i: (keys[i], i)), dtype=np.int64)    # stable sort by key
    sa = order.copy()
    skeys = keys[order]
    unit = sum(radix ** j for j in range(k))
    allc = {c * unit for c in range(1, sigma + 1)}

    # maximal runs of >= k equal symbols
    runs = []
    i = 0
    while i < n:
        j = i
        while j + 1 < n and text[j + 1] == text[i]:
            j += 1
        if j - i + 1 >= k:
            end = j + 1
            runs.append((i, end, end < n and text[end] > text[i]))
        i = j + 1

    def run_remaining(s):
        for start, end, larger in runs:
            if start <= s < end and end - s >= k:
                return end - s, larger
        return 0, False

    rank = np.full(n, -1, np.int64)                                          # -1 = never stored
    # groups of equal keys
    bnd = np.ones(n + 1, bool)
    bnd[1:n] = skeys[1:] != skeys[:-1]
    starts = np.flatnonzero(bnd[:n])
    glen = np.diff(np.append(starts, n))
    general, run_list = [], []                                               # lists of (slot, suffix, gid)
    gid = 0
    for st, ln in zip(starts, glen):
        if ln == 1:
            continue
        is_run = use_runs and skeys[st] in allc and len(runs) > 0
        for t in range(st, st + ln):
            rank[sa[t]] = st
        if is_run:
            run_list += [(t, int(sa[t])) for t in range(st, st + ln)]
        else:
            general += [(t, int(sa[t]), gid) for t in range(st, st + ln)]
            gid += 1

    def regroup(entries, sort_key, gid_base):
        """entries: list of (slot, suffix) in slot order, already keyed by sort_key(index) -> comparable.  Sort stably,
        split into groups of equal keys, finish singletons, return the kept (slot, suffix, gid) list."""
        slots = [e[0] for e in entries]
        ks = [sort_key(u) for u in range(len(entries))]                      # all keys exist before any rank changes
        idx = sorted(range(len(entries)), key=lambda u: (ks[u], u))
        kept, g = [], gid_base
        u = 0
        while u < len(idx):
            v = u
            while v + 1 < len(idx) and ks[idx[v + 1]] == ks[idx[u]]:
                v += 1
            for w in range(u, v + 1):
                s = entries[idx[w]][1]
                rank[s] = slots[u]
                if u == v:
                    sa[slots[w]] = s
                else:
                    kept.append((slots[w], s, g))
            if v > u:
                g += 1
            u = v + 1
        return kept, g

    lst = general
    groups = gid
    if run_list:                                                             # stage A: one sort of the long-run suffixes
        def run_key(u):
            s = run_list[u][1]
            r, larger = run_remaining(s)
            assert r >= k
            return (int(code_of[text[s]]), 1, -r) if larger else (int(code_of[text[s]]), 0, r)
        kept, groups = regroup(run_list, run_key, groups)
        lst = lst + kept
    h = k
    rounds = 0
    def slot_of(q):
        # lower bound of key(q) in the sorted keys (the key is unique: the suffix was final after the first sort)
        kq = keys[q]
        lo, hi = 0, n
        while lo < hi:
            mid = (lo + hi) // 2
            if skeys[mid] < kq:
                lo = mid + 1
            else:
                hi = mid
        return lo

    while lst:
        assert h < 2 * n + 64, "did not converge"

        def key(u):
            slot, s, g = lst[u]
            q = s + h
            r, _ = run_remaining(s) if use_runs and runs else (0, False)
            if r:
                q += r - k
            if q >= n:
                return (g, 0)
            rk = rank[q]
            if rk < 0:                                                    # final since the first sort: its slot
                rk = slot_of(q)
            return (g, int(rk) + 1)

        entries = [(slot, s) for slot, s, _ in lst]
        lst, groups = regroup(entries, key, 0)
        h *= 2
        rounds += 1
    return sa
