#!/usr/bin/env python
"""variants.py — check and time every experimental kernel variant in ONE GPU call.

    python tools/variants.py [--length 248956422] [--steps 5] [--out gpurun_out/variants.jsonl] [--only NAME ...]

The variants are selected by environment variables read once per process, so each one runs in a child process:
  1. parity: a block of every "family" below is built and compared byte for byte with the build of the DEFAULT kernels
     (which the GPU parity tests compare with the oracle) — suffix array, BWT, .gcz body, .gcx body;
  2. time: --steps device-resident builds of the chr1-shaped block; mean of total / sort / wavelet phase times and of the
     digit passes (gcz_last_build_timing).
One JSON line per variant.  Needs a GPU; nothing here is a bench value (no clock sampling, no e2e).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

VARIANTS = {
    "default": {},
    "sort8_dst32": {"GCZ_SORT_VARIANT": "8"},
    "sort9_early_lookback": {"GCZ_SORT_VARIANT": "9"},
    "sort10_both": {"GCZ_SORT_VARIANT": "10"},
    "sort11_9bit": {"GCZ_SORT_VARIANT": "11"},
    "sort12_9bit_both": {"GCZ_SORT_VARIANT": "12"},
    "sort13_prefetch_values": {"GCZ_SORT_VARIANT": "13"},
    "sort14_9bit_prefetch": {"GCZ_SORT_VARIANT": "14"},
    "emit_lut": {"GCZ_EMIT_VARIANT": "1"},
    "text_hist_uniform": {"GCZ_TEXT_HIST_VARIANT": "1"},
    "bwt_packed": {"GCZ_BWT_VARIANT": "1"},
    "count_sorted": {"GCZ_COUNT_SORT": "1"},        # queries only: digests + count_ms below
    "locate_refill": {"GCZ_LOCATE_VARIANT": "1"},   # queries only: the digest check here; time it with tools/run_configs.py (cfg5)
    "early_marker": {"GCZ_EARLY_MARKER": "1"},      # host outputs only: parity here, timing through `GCZ_EARLY_MARKER=1 python bench.py` (e2e)
    "all_small": {"GCZ_EMIT_VARIANT": "1", "GCZ_TEXT_HIST_VARIANT": "1", "GCZ_BWT_VARIANT": "1"},
    "all_small_9bit": {"GCZ_EMIT_VARIANT": "1", "GCZ_TEXT_HIST_VARIANT": "1", "GCZ_BWT_VARIANT": "1", "GCZ_SORT_VARIANT": "11"},
}


def families():
    import numpy as np
    from gecoz_b200 import synth
    rng = np.random.default_rng(11)
    acgtn = np.frombuffer(b"ACGTN", np.uint8)
    yield "tiny", np.frombuffer(b"GATTACA\0", np.uint8).copy()
    yield "two_strings", np.frombuffer(b"ACGTN\0ACG\0", np.uint8).copy()
    yield "iid_300k", synth.cfg1_text(300_000)
    yield "chr_shaped_3M", synth.cfg2_text(3_000_000)
    yield "all_same", synth.block_of([np.full(50_000, ord("A"), np.uint8)])
    yield "tandem", synth.block_of([np.frombuffer(b"ACACACACGT" * 30_000, np.uint8)])
    yield "multi", synth.block_of([synth.iid_acgtn(30_000, 5), synth.iid_acgtn(20_000, 6), synth.iid_acgtn(7, 7),
                                   np.zeros(0, np.uint8), synth.iid_acgtn(20_000, 6)])
    yield "lower_iupac", synth.block_of([np.frombuffer(b"ACGTNacgtnRYKM", np.uint8)[rng.integers(0, 14, 150_000)]])
    lens = rng.integers(1, 200, 4000)
    yield "runs_mixed", synth.block_of([np.repeat(acgtn[rng.integers(0, 5, len(lens))], lens)])
    yield "acgt_only", synth.block_of([np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 1_000_003)]])
    yield "eight_symbols", synth.block_of([np.frombuffer(b"ACGTNRY", np.uint8)[rng.integers(0, 7, 400_001)]])


def build(G, text, rate=32):
    import numpy as np
    shape = G.shape_from_counts(G.symbol_counts(text))
    gcz = np.zeros(int(shape.size), np.uint8)
    gcx = np.zeros(G.index_size(len(text), rate.bit_length() - 1), np.uint8)
    sa = np.zeros(len(text), np.int32)
    bwt = np.zeros(len(text), np.uint8)
    t = G.build_block(0, text, len(text), rate, shape, gcz, gcx, sa, bwt)
    return gcz, gcx, sa, bwt, t


def child(args) -> None:
    """Runs under the variant's environment: builds every family, stores or compares the results, then times cfg2."""
    import hashlib
    import numpy as np
    import torch
    import gecoz_b200 as G
    from gecoz_b200 import synth
    G.lib()
    ref_path = Path(args.ref)
    ref = json.loads(ref_path.read_text()) if ref_path.exists() else None
    digests, mismatches = {}, []
    for name, text in families():
        for rate in (32, 4):
            gcz, gcx, sa, bwt, _ = build(G, text, rate)
            d = [hashlib.sha256(a.tobytes()).hexdigest() for a in (gcz, gcx, sa, bwt)]
            digests[f"{name}/{rate}"] = d
            if ref is not None and ref.get(f"{name}/{rate}") != d:
                what = [w for w, x, y in zip(("gcz", "gcx", "sa", "bwt"), ref.get(f"{name}/{rate}", [None] * 4), d) if x != y]
                mismatches.append(f"{name}/{rate}: {','.join(what)}")
    # the sorter on its own: pairs and keys only (the query path sorts keys only), odd sizes and bit widths
    from gecoz_b200 import _native as N
    rng = np.random.default_rng(7)
    for n_keys, bits, with_vals in ((1, 64, True), (6143, 64, True), (6145, 64, False), (1_000_003, 64, True), (2_000_001, 45, False),
                                    (300_000, 41, True), (250_000, 9, False), (250_000, 8, True), (3_000_000, 63, False)):
        keys = rng.integers(0, 2 ** 63, n_keys, dtype=np.uint64)
        if bits < 64:
            keys &= np.uint64((1 << bits) - 1)
        if n_keys > 1000:
            keys[: n_keys // 3] = keys[n_keys // 3: 2 * (n_keys // 3)]
        vals = np.arange(n_keys, dtype=np.uint32)
        k2, v2 = keys.copy(), vals.copy()
        N.check(G.lib().gcz_dbg_sort_pairs(0, N.ptr(k2), N.ptr(v2) if with_vals else None, n_keys, 0, bits))
        order = np.argsort(keys, kind="stable")
        if not np.array_equal(k2, keys[order]) or (with_vals and not np.array_equal(v2, vals[order])):
            mismatches.append(f"sort n={n_keys} bits={bits} vals={with_vals}")
    # queries on a merged block (the locate post-processing sorts keys only): intervals, positions, per-string split
    seqs = [synth.iid_acgtn(500_000, 31), synth.iid_acgtn(300_000, 32), synth.iid_acgtn(7, 33)]
    qtext = synth.block_of(seqs)
    gcz, gcx, _, _, _ = build(G, qtext, 32)
    g = G.GSSA.open(0, gcz, len(qtext), gcx)
    data, off = synth.patterns(qtext, 50_000, 6, 40, seed=8)
    sp, ep = g.count_batch(packed=(data, off))
    per, pos, poff = g.find_batch_raw(packed=(data, off))
    g.close()
    qd = [hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() for a in (sp, ep, per, pos, poff)]
    digests["queries"] = qd
    if ref is not None and ref.get("queries") != qd:
        mismatches.append("queries: " + ",".join(w for w, x, y in zip(("sp", "ep", "per", "pos", "off"), ref.get("queries", [None] * 5), qd) if x != y))
    if ref is None:
        ref_path.write_text(json.dumps(digests))
    # timing: device-resident builds of the chr1-shaped block
    text = synth.cfg2_text(args.length, seed=3)
    n = len(text)
    d_text = torch.from_numpy(text).cuda()
    shape = G.shape_from_counts(G.symbol_counts(d_text, 0))
    d_gcz = torch.empty(int(shape.size), dtype=torch.uint8, device="cuda")
    d_gcx = torch.empty(G.index_size(n, 5), dtype=torch.uint8, device="cuda")
    infos = []
    for i in range(2 + args.steps):
        t = G.build_block(0, d_text, n, 32, shape, d_gcz, d_gcx)
        if i >= 2:
            infos.append(t)
    keys = ("total_ms", "sort_initial_ms", "sort_refine_ms", "bwt_hswt_ms", "ssa_ms", "radix_ms", "radix_full_ms", "radix_text_ms",
            "radix_launches", "radix_full_launches", "kernel_launches")
    mean = {k: float(np.mean([t[k] for t in infos])) for k in keys}
    big = hashlib.sha256(d_gcz.cpu().numpy().tobytes()).hexdigest()[:16] + hashlib.sha256(d_gcx.cpu().numpy().tobytes()).hexdigest()[:16]
    # count and find on that index: 1 M patterns resident on the device (wall clock around the call, best of three)
    import time
    g = G.GSSA.open(0, d_gcz, n, d_gcx)
    pdata, poff = synth.patterns(text, 1_000_000, 15, 100, seed=5)
    dp, do = torch.from_numpy(pdata).cuda(), torch.from_numpy(poff).cuda()
    d_sp = torch.empty(len(poff) - 1, dtype=torch.int64, device="cuda")
    d_ep = torch.empty_like(d_sp)
    count_ms, find_ms = [], []
    for _ in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g.count_batch(packed=(dp, do), out=(d_sp, d_ep))
        torch.cuda.synchronize()
        count_ms.append((time.perf_counter() - t0) * 1e3)
    small = (pdata[:int(poff[100_000])], poff[:100_001])
    for _ in range(3):
        t0 = time.perf_counter()
        g.find_batch_raw(packed=small)
        find_ms.append((time.perf_counter() - t0) * 1e3)
    g.close()
    mean["count_1M_ms"], mean["find_100k_ms"] = float(min(count_ms[1:])), float(min(find_ms[1:]))
    print("RESULT " + json.dumps({"variant": args.child, "env": VARIANTS[args.child], "parity_vs_default": ("stored" if ref is None else "ok") if not mismatches
                                  else mismatches, "cfg2_digest": big, "n": n, **mean}), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--length", type=int, default=248_956_422)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "variants.jsonl"))
    ap.add_argument("--only", nargs="*", default=None)
    ap.add_argument("--child", default=None)
    ap.add_argument("--ref", default="/tmp/gcz_variants_ref.json")
    args = ap.parse_args()
    if args.child:
        child(args)
        return
    out = Path(args.out)
    out.parent.mkdir(parents=True, exist_ok=True)
    Path(args.ref).unlink(missing_ok=True)
    names = ["default"] + [v for v in VARIANTS if v != "default" and (not args.only or v in args.only)]
    base_digest = None
    with open(out, "w") as f:
        for name in names:
            env = dict(os.environ, **VARIANTS[name])
            try:                                  # a variant that hangs (a look-back that never resolves) must not hold the box
                r = subprocess.run([sys.executable, __file__, "--child", name, "--length", str(args.length), "--steps", str(args.steps),
                                    "--ref", args.ref], env=env, capture_output=True, text=True, timeout=240)
                line = next((l[7:] for l in r.stdout.splitlines() if l.startswith("RESULT ")), None)
                rec = json.loads(line) if line else {"variant": name, "env": VARIANTS[name], "failed": r.returncode,
                                                     "stderr": r.stderr[-1500:]}
            except subprocess.TimeoutExpired:
                line, rec = None, {"variant": name, "env": VARIANTS[name], "failed": "timeout after 240 s"}
            if name == "default" and line:
                base_digest = rec["cfg2_digest"]
            elif line:
                rec["cfg2_same_as_default"] = rec["cfg2_digest"] == base_digest
            f.write(json.dumps(rec) + "\n")
            f.flush()
            print(json.dumps({k: rec.get(k) for k in ("variant", "parity_vs_default", "cfg2_same_as_default", "total_ms", "sort_initial_ms",
                                                      "bwt_hswt_ms", "radix_full_ms", "radix_text_ms", "count_1M_ms", "find_100k_ms", "failed")}), flush=True)


if __name__ == "__main__":
    main()
