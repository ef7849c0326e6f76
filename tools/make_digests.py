#!/usr/bin/env python
"""make_digests.py — golden digests of the BASELINE.json configs at FULL size, computed by the CPU oracle alone.

    python tools/make_digests.py [--out tests/golden/full_size_digests.json] [--workers 4] [--only cfg1 cfg2 cfg3 queries]

Test infrastructure (it runs oracle/ and nothing of the product): the GPU parity tests and bench.py compare the sha256 of
what the CUDA path produces with these, so byte parity at the benchmarked sizes costs no oracle run on the GPU box
(the oracle needs ~25 s per chr1-sized block).  Needs ~2 minutes on 8 cores and < 8 GB here.

  cfg1   16 000 001-symbol block:  text, SA (int32 LE), BWT, .gcz body, .gcx body
  cfg2   248 956 423-symbol block: the same five
  cfg3   the 18 blocks tools/GecoIndex.java:72-98 makes of the 25 hg38-length sequences (oracle merge): per block the
         member headers and the same five digests; the two whole files (headers + bodies in file order)
  queries (cfg4 / cfg5 at full size on the blocks where per-string results differ from intervals: the merged blocks
         chr13+chr14 and chr15+chr22+chr21+chrM, and the chr11 block of the `-s chr11` filter): 100 000 patterns of
         length 15..100 each — sha256 of sp[], ep[] (GSSA.search) and of the per-string counts / positions of GSSA.find
         for the first 20 000 of them
"""
from __future__ import annotations

import argparse
import hashlib
import json
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

QUERY_BLOCKS = ("chr13", "chr15", "chr11")          # first member of the blocks whose query results are pinned
QUERY_PATTERNS = 100_000
FIND_PATTERNS = 20_000


def sha(a) -> str:
    return hashlib.sha256(memoryview(np.ascontiguousarray(a)).cast("B")).hexdigest()


def block_digests(O, text: np.ndarray, keep_bodies: bool = False) -> dict:
    r = O.build_block(text, 32, want_sa=True, want_bwt=True, threads=2)
    d = {"n": int(len(text)), "text": sha(text), "sa": sha(r["sa"]), "bwt": sha(r["bwt"]),
         "gcz_body": sha(r["gcz_body"]), "gcx_body": sha(r["gcx_body"]),
         "gcz_body_len": int(len(r["gcz_body"])), "gcx_body_len": int(len(r["gcx_body"]))}
    if keep_bodies:
        d["_bodies"] = (r["gcz_body"], r["gcx_body"])
    return d


def query_digests(O, synth, text: np.ndarray, gcz: np.ndarray, gcx: np.ndarray, seed: int) -> dict:
    g = O.GSSA(gcz, len(text), gcx)
    data, off = synth.patterns(text, QUERY_PATTERNS, 15, 100, seed=seed)
    sp, ep, calls = g.search_batch(data, off)
    per = np.zeros((FIND_PATTERNS, g.n_strings), np.int64)
    positions = []
    for q in range(FIND_PATTERNS):
        res = g.find(data[off[q]:off[q + 1]].tobytes())
        if res is None:
            continue
        for s, arr in enumerate(res):
            if arr is not None:
                per[q, s] = len(arr)
                positions.append(arr)
    pos = np.concatenate(positions) if positions else np.zeros(0, np.int64)
    out = {"seed": seed, "patterns": QUERY_PATTERNS, "find_patterns": FIND_PATTERNS, "n_strings": int(g.n_strings),
           "pattern_bytes": sha(data), "sp": sha(sp), "ep": sha(ep), "found": int((ep >= sp).sum()), "rank_calls": int(calls),
           "per_string_counts": sha(per), "positions": sha(pos), "occurrences": int(len(pos)),
           "string_ends": [int(x) for x in g.string_ends()]}
    g.close()
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden" / "full_size_digests.json"))
    ap.add_argument("--workers", type=int, default=4)
    ap.add_argument("--only", nargs="*", default=["cfg1", "cfg2", "cfg3", "queries"])
    args = ap.parse_args()

    from gecoz_b200 import synth            # the generators only (numpy); nothing of the CUDA path is touched
    from oracle import gcz_oracle as O
    O.build()
    out_path = Path(args.out)
    gold = json.loads(out_path.read_text()) if out_path.exists() else {}
    gold["how"] = "python tools/make_digests.py — oracle/ (C restatement of the Java path) on gecoz_b200.synth inputs; sha256"
    gold["numpy"] = np.__version__
    t0 = time.time()

    if "cfg1" in args.only:
        gold["cfg1"] = block_digests(O, synth.cfg1_text())
        print(f"cfg1 done {time.time() - t0:.0f}s", flush=True)
    if "cfg2" in args.only:
        gold["cfg2"] = block_digests(O, synth.cfg2_text())
        print(f"cfg2 done {time.time() - t0:.0f}s", flush=True)

    if "cfg3" in args.only or "queries" in args.only:
        names, lengths = synth.HG38_NAMES, synth.HG38_LENGTHS
        blocks = O.merge_blocks(lengths, names)

        def one(b):
            ids = blocks[b]
            text = synth.block_of([synth.chromosome_shaped(lengths[i], 4 + i) for i in ids])
            want_q = "queries" in args.only and names[ids[0]] in QUERY_BLOCKS
            d = block_digests(O, text, keep_bodies=True)
            gcz, gcx = d.pop("_bodies")
            hs = [names[i] for i in ids]
            d["headers"] = hs
            hdr_len = 26 + sum(len(h) + 1 for h in hs)
            d["_ref_header"] = O.ref_header(hs, hdr_len + len(gcz), len(text))
            d["_ssa_header"] = O.ssa_header(hs, len(gcx))
            if want_q:
                d["queries"] = query_digests(O, synth, text, gcz, gcx, seed=500 + b)
            # whole-file digests need the bodies in file order: keep them only as long as needed
            d["_gcz"], d["_gcx"] = gcz, gcx
            print(f"cfg3 block {b} {hs} done {time.time() - t0:.0f}s", flush=True)
            return d

        hz, hx = hashlib.sha256(), hashlib.sha256()
        zlen = xlen = 0
        per_block = [None] * len(blocks)
        with ThreadPoolExecutor(args.workers) as pool:
            futs = [pool.submit(one, b) for b in range(len(blocks))]
            for b, f in enumerate(futs):                      # file order
                d = f.result()
                for h, parts in ((hz, (d.pop("_ref_header"), d.pop("_gcz"))), (hx, (d.pop("_ssa_header"), d.pop("_gcx")))):
                    for p in parts:
                        h.update(memoryview(np.ascontiguousarray(np.frombuffer(p, np.uint8) if isinstance(p, bytes) else p)).cast("B"))
                zlen += 26 + sum(len(h) + 1 for h in d["headers"]) + d["gcz_body_len"]
                xlen += 25 + d["gcx_body_len"]
                per_block[b] = d
        if "cfg3" in args.only:
            gold["cfg3"] = {"blocks": per_block, "gcz_file": hz.hexdigest(), "gcx_file": hx.hexdigest(),
                            "gcz_file_len": zlen, "gcx_file_len": xlen,
                            "bases": int(sum(lengths)), "symbols": int(sum(lengths) + len(lengths))}
        else:
            for b, d in enumerate(per_block):
                if "queries" in d:
                    gold["cfg3"]["blocks"][b]["queries"] = d["queries"]
        print(f"cfg3 done {time.time() - t0:.0f}s", flush=True)

    out_path.parent.mkdir(parents=True, exist_ok=True)
    out_path.write_text(json.dumps(gold, indent=1) + "\n")
    print(f"wrote {out_path} in {time.time() - t0:.0f}s")


if __name__ == "__main__":
    main()
