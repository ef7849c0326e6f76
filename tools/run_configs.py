#!/usr/bin/env python
"""BASELINE.json configs 3-5 at full size (or scaled down with --scale): hg38-shaped multi-block build, batched count
and batched locate against the resulting multi-block index.  Prints one JSON line per config (rank 0).

    python tools/run_configs.py [--scale 1.0] [--count-patterns 100000000] [--locate-patterns 10000000] [--out DIR]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_configs.py ...

cfg3  25 sequences with the hg38 lengths -> 18 chromosome-bounded blocks (tools/GecoIndex.java:72-98), LPT-sharded over
      the ranks, one .gcz/.gcx pair (gecoz_b200/sharding.py).  Reported: wall-clock Mbp/s of the whole call (host block
      assembly, H2D, build, D2H into the mmap'd file) and the Mbp/s of the device time alone; sha256 of both files
      (identical for every N: the sharded build is bit-exact with the single-rank one).
cfg4  count: patterns of length 15..100 (half sampled N-free from the genome) in chunks, every chunk against every
      block of the replicated index, shards of a chunk gathered on rank 0.
cfg5  locate: GSSA.find of every pattern against every block (per-string positions), then the `-s chr11` form:
      only the block holding chr11, only its string.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def sha256(path) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="sequence lengths = hg38 lengths x scale")
    ap.add_argument("--count-patterns", type=int, default=100_000_000)
    ap.add_argument("--locate-patterns", type=int, default=10_000_000)
    ap.add_argument("--chunk", type=int, default=4_000_000, help="patterns generated / searched per chunk")
    ap.add_argument("--out", default="/tmp/gecoz_cfg3")
    ap.add_argument("--skip-queries", action="store_true")
    args = ap.parse_args()

    import torch
    import gecoz_b200 as G
    from gecoz_b200 import sharding, synth

    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(dev))
    G.lib()
    out = Path(args.out)
    if rank == 0:
        out.mkdir(parents=True, exist_ok=True)

    def emit(d: dict) -> None:
        if rank == 0:
            print(json.dumps(d), flush=True)

    # ---- cfg3: sharded build ---------------------------------------------------------------------------------------
    # sequences are synthesised on demand, so that a rank only pays for the blocks it builds (eight GPUs must not wait
    # for eight copies of the same host work)
    class LazySequence:
        def __init__(self, length, seed):
            self.length, self.seed, self._data = length, seed, None

        def __len__(self):
            return self.length

        def __call__(self):
            if self._data is None:
                self._data = synth.chromosome_shaped(self.length, self.seed)
            return self._data

    t0 = time.perf_counter()
    recs = [(name, LazySequence(max(8, int(length * args.scale)), 4 + i))
            for i, (name, length) in enumerate(zip(synth.HG38_NAMES, synth.HG38_LENGTHS))]
    from gecoz_b200.geco_index import FastaSequence, merge_blocks
    plan = merge_blocks([FastaSequence(h, len(s), None, i) for i, (h, s) in enumerate(recs)])
    for b, r in zip(plan, sharding.lpt_assign([b.size for b in plan], world)):
        if r == rank:
            for seq in b.sequences:
                recs[seq.id][1]()                         # my blocks' sequences exist before the clock starts
    gen_s = time.perf_counter() - t0
    bases = sum(len(s) for _, s in recs)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    info = sharding.sharded_index_records(recs, out / "hg38s.gcz", rank=rank, world=world, engine=sharding.GpuEngine(local_rank))
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    dev_ms = sum(t["total_ms"] for t in info["timings"])
    if world > 1:
        tt = torch.tensor([wall, dev_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        wall, dev_ms = (float(x) for x in tt.tolist())
    emit({"config": "cfg3", "metric": "FM-index build throughput, hg38-shaped multi-block file", "n_gpus": world, "scale": args.scale,
          "bases": bases, "blocks": len(info["blocks"]), "blocks_per_rank": [info["owner"].count(r) for r in range(world)],
          "wall_s": wall, "value": bases / 1e6 / wall, "unit": "Mbp/s (wall clock of the whole sharded call, max over ranks)",
          "device_ms_max_rank": dev_ms, "device_only_mbp_s": bases / 1e6 / (dev_ms / 1e3),
          "gcz_bytes": (out / "hg38s.gcz").stat().st_size, "gcx_bytes": (out / "hg38s.gcx").stat().st_size,
          "gcz_sha256": sha256(out / "hg38s.gcz") if rank == 0 else None, "gcx_sha256": sha256(out / "hg38s.gcx") if rank == 0 else None,
          "synthesis_s": gen_s, "data": "synthetic"})
    if args.skip_queries:
        return

    # ---- replicated index ---------------------------------------------------------------------------------------------
    t0 = time.perf_counter()
    reader = G.GecozFileReader(out / "hg38s.gcz", local_rank)
    bheaders = reader.getBlockHeaders()
    gssas = [reader.read(h) for h in bheaders]
    torch.cuda.synchronize()
    open_s = time.perf_counter() - t0
    by_header = {h: s for h, s in recs}

    # A few distinct chunks, drawn once on rank 0 and broadcast, are cycled through: pattern synthesis on the host is
    # slower than the searches and must not run on every rank while the GPUs are held.
    distinct = {}

    def pattern_chunk(i: int, count: int):
        key = (i % 3, count)
        if key not in distinct:
            if rank == 0:
                name = synth.HG38_NAMES[(i % 3) * 7 % 24]
                data, off = synth.patterns(by_header[name](), count, 15, 100, seed=5000 + i % 3)
            if world > 1:
                size = torch.tensor([len(data) if rank == 0 else 0], dtype=torch.int64, device=dev)
                dist.broadcast(size, 0)
                t_data = torch.from_numpy(data).to(dev) if rank == 0 else torch.empty(int(size.item()), dtype=torch.uint8, device=dev)
                t_off = torch.from_numpy(off).to(dev) if rank == 0 else torch.empty(count + 1, dtype=torch.int64, device=dev)
                dist.broadcast(t_data, 0)
                dist.broadcast(t_off, 0)
                data, off = t_data.cpu().numpy(), t_off.cpu().numpy()
            distinct[key] = (data, off)
        return distinct[key]

    # ---- cfg4: count ---------------------------------------------------------------------------------------------------
    done, found, ms = 0, 0, 0.0
    i = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    while done < args.count_patterns:
        cnt = min(args.chunk, args.count_patterns - done)
        data, off = pattern_chunk(i, cnt)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = sharding.count_totals_sharded(gssas, data, off, rank=rank, world=world, device=dev)
        torch.cuda.synchronize()
        ms += (time.perf_counter() - t0) * 1e3
        if rank == 0:
            found += int((res > 0).sum())
        done += cnt
        i += 1
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    emit({"config": "cfg4", "metric": "count queries/s against the hg38-shaped index (every pattern x every block)", "n_gpus": world,
          "patterns": done, "blocks": len(gssas), "patterns_found_somewhere": found, "ms": ms, "value": done / (ms / 1e3),
          "unit": "queries/s (wall clock of count_totals_sharded: H2D of the shard from pageable memory, one search per block, per-pattern totals back, gather; host pattern synthesis excluded)",
          "index_open_s": open_s, "data": "synthetic"})

    # ---- cfg5: locate ---------------------------------------------------------------------------------------------------
    done, occ, ms = 0, 0, 0.0
    i = 0
    while done < args.locate_patterns:
        cnt = min(args.chunk, args.locate_patterns - done)
        data, off = pattern_chunk(1000 + i, cnt)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = sharding.find_sharded(gssas, data, off, rank=rank, world=world, device=dev)
        torch.cuda.synchronize()
        ms += (time.perf_counter() - t0) * 1e3
        if rank == 0:
            occ += int(sum(len(r[1]) for r in res))
        done += cnt
        i += 1
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    emit({"config": "cfg5", "metric": "locate (GSSA.find: per-string sorted positions) against the hg38-shaped index", "n_gpus": world,
          "patterns": done, "occurrences": occ, "ms": ms, "value": done / (ms / 1e3), "unit": "patterns/s (wall clock of find_sharded)",
          "occurrences_per_s": occ / (ms / 1e3), "data": "synthetic"})

    # `-s chr11 PATTERN` for a batch: only the block that holds chr11, only that string
    bh = reader.findBlockHeader("chr11")
    g11 = gssas[bheaders.index(bh)]
    nstr = bh.findHeader("chr11")
    cnt = min(args.chunk, args.locate_patterns)
    data, off = synth.patterns(by_header["chr11"](), cnt, 15, 100, seed=77)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    per, pos, poff = g11.find_batch_raw(packed=(data, off))
    sec = time.perf_counter() - t0
    emit({"config": "cfg5-filter", "metric": "locate with the header filter chr11 (one block, one string reported)", "n_gpus": 1,
          "patterns": cnt, "occurrences_in_chr11": int(per[:, nstr].sum()), "ms": sec * 1e3, "value": cnt / sec, "unit": "patterns/s",
          "data": "synthetic"})
    for g in gssas:
        g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
