"""Multi-GPU partitioning of the FM-index path: one process per GPU, `torch.distributed` for the plumbing.

The reference has no communication layer: its only parallelism is block-level tasks on a JDK pool
(fmt/GecozFileWriter.java:174-227).  The same two decompositions carry over to a multi-GPU box:

* **build** — blocks share nothing and their file offsets are fixed before they are computed
  (fmt/GecozFileWriter.java:135-156), so block -> rank by longest-processing-time-first and every rank writes
  its own slices of the shared `.gcz` / `.gcx`.  No data-path collective; the ranks only exchange the byte
  size of each block's wavelet tree (control metadata, one all_gather_object) so that all of them can compute
  the offsets.
* **query** — patterns are independent and every pattern visits every block (tools/GecoMatch.java:114-131), so
  the index is replicated, the batch is cut into contiguous equal shards and the results come back to rank 0
  with one gather per result array (fixed-size for intervals and per-string counts, length-prefixed for the
  located positions).

`group`/backend: NCCL on the GPUs (tensors staged on `device`), gloo on CPU for the host-logic tests.
"""
from __future__ import annotations

import os
import time
from pathlib import Path
from typing import Iterable, Sequence

import numpy as np

from . import _native as N
from .gecoz_file import (GecozRefBlockHeader, GecozSSABlockHeader, _BodyWriter, _block_layout, build_block, index_size,
                         shape_from_counts, ssa_path_for, symbol_counts)
from .geco_index import FastaSequence, block_text, merge_blocks


# ---- partitioning (pure functions) ---------------------------------------------------------------------------
def lpt_assign(sizes: Sequence[int], world: int) -> list[int]:
    """Longest-processing-time-first: blocks by decreasing size (ties: file order) onto the least loaded rank
    (ties: lowest rank).  Build cost is linear in the block size to a good approximation."""
    load = [0] * world
    owner = [0] * len(sizes)
    for i in sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i)):
        r = min(range(world), key=lambda r: (load[r], r))
        owner[i] = r
        load[r] += int(sizes[i])
    return owner


def shard_bounds(n_items: int, world: int) -> list[tuple[int, int]]:
    """Contiguous shards whose sizes differ by at most one, in rank order."""
    base, rem = divmod(n_items, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


# ---- collectives ---------------------------------------------------------------------------------------------------
def _dist():
    import torch.distributed as dist
    return dist


def _barrier(world, group):
    if world > 1:
        _dist().barrier(group=group)


def gather_varlen(arr: np.ndarray, *, rank: int, world: int, group=None, device=None) -> list[np.ndarray] | None:
    """Gather one int64 array per rank to rank 0 (lengths first, then one padded gather).  Returns the list on
    rank 0 and None elsewhere."""
    arr = np.ascontiguousarray(arr, dtype=np.int64).reshape(-1)
    if world == 1:
        return [arr]
    import torch
    dist = _dist()
    dev = torch.device(device) if device is not None else torch.device("cpu")
    mine = torch.tensor([arr.size], dtype=torch.int64, device=dev)
    each = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(each, mine, group=group)
    lens = torch.cat(each)
    lens_h = lens.cpu().numpy()
    width = int(lens_h.max())
    if width == 0:
        return [np.zeros(0, np.int64) for _ in range(world)] if rank == 0 else None
    padded = torch.zeros(width, dtype=torch.int64, device=dev)
    padded[:arr.size] = torch.from_numpy(arr).to(dev)
    parts = [torch.empty(width, dtype=torch.int64, device=dev) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, parts, dst=0, group=group)
    if rank != 0:
        return None
    return [p[:int(n)].cpu().numpy() for p, n in zip(parts, lens_h)]


# ---- build -----------------------------------------------------------------------------------------------------------
class GpuEngine:
    """The CUDA path behind the C ABI on one device (the only engine the product uses)."""

    def __init__(self, device: int = 0):
        self.device = device

    def symbol_counts(self, text: np.ndarray) -> np.ndarray:
        return symbol_counts(text, self.device)

    def build_block(self, text, n, sampling_rate, shape, gcz_out, gcx_out) -> dict:
        return build_block(self.device, text, n, sampling_rate, shape, gcz_out, gcx_out)


def sharded_index_records(records: Iterable[tuple[str, object]], opath, xpath=None, sampling: int = 32, *,
                          rank: int = 0, world: int = 1, engine=None, group=None) -> dict:
    """GecoIndex.index (tools/GecoIndex.java:51-117) across `world` ranks writing ONE `.gcz`/`.gcx` pair.

    Every rank runs the same deterministic block merge, takes its LPT share of the blocks, and writes the
    headers and bodies of those blocks at offsets all ranks agree on.  The files are byte-identical to a
    single-process build.  `opath` must be on a file system all ranks see (one node: always true)."""
    engine = engine or GpuEngine(int(os.environ.get("LOCAL_RANK", "0")))
    opath = Path(opath)
    xpath = Path(xpath) if xpath is not None else ssa_path_for(opath)
    sf = sampling.bit_length() - 1
    t0 = time.perf_counter()

    seqs = []
    for i, (h, s) in enumerate(records):
        seqs.append(FastaSequence(h, s.length if hasattr(s, "length") else len(s), s, i))
    blocks = merge_blocks(seqs)
    if not blocks:
        raise ValueError("no data found")
    owner = lpt_assign([b.size for b in blocks], world)
    mine = [i for i, r in enumerate(owner) if r == rank]

    # my blocks: text, histogram, shape (the tree's byte size is the one thing the other ranks cannot know)
    prepared, sizes = {}, {}
    for i in mine:
        headers, text = block_text(blocks[i])
        shape = shape_from_counts(engine.symbol_counts(text))
        prepared[i] = (headers, text, shape)
        sizes[i] = int(shape.size)
    if world > 1:
        every = [None] * world
        _dist().all_gather_object(every, sizes, group=group)
        sizes = {k: v for d in every for k, v in d.items()}

    ref_off, ssa_off, ref_pos, ssa_pos = [], [], 0, 0
    for i, b in enumerate(blocks):                                   # file order == block order
        headers = [s.header for s in b.sequences]
        n = sum(s.length + 1 for s in b.sequences)
        ref_off.append(ref_pos)
        ssa_off.append(ssa_pos)
        ref_pos += GecozRefBlockHeader.block_header_length(headers) + sizes[i]
        ssa_pos += GecozSSABlockHeader.LENGTH + index_size(n, sf)

    if rank == 0:
        for p, size in ((opath, ref_pos), (xpath, ssa_pos)):
            with open(p, "w+b") as f:
                os.ftruncate(f.fileno(), size)
    _barrier(world, group)

    timings = []
    # block k + 1 is uploaded (re-counted: that is what stages its text on the device) while block k is being built, and
    # the bodies of block k - 1 are on their way to the files (reused host buffers, pwrite on helper threads)
    from concurrent.futures import ThreadPoolExecutor
    bodies = _BodyWriter(max_buffers=3)
    with open(opath, "r+b") as fref, open(xpath, "r+b") as fssa, ThreadPoolExecutor(1) as stager:
        try:
            staged = stager.submit(engine.symbol_counts, prepared[mine[0]][1]) if mine else None
            for k, i in enumerate(mine):
                headers, text, shape = prepared.pop(i)
                staged.result()
                if k + 1 < len(mine):
                    staged = stager.submit(engine.symbol_counts, prepared[mine[k + 1]][1])
                n = len(text)
                idx_size = index_size(n, sf)
                hdr = GecozRefBlockHeader(headers, GecozRefBlockHeader.block_header_length(headers) + int(shape.size), n)
                hb, sb = hdr.to_bytes(), GecozSSABlockHeader(headers, idx_size).to_bytes()
                ref_at, ssa_at, size = _block_layout(len(hb), int(shape.size), idx_size)
                buf = bodies.take(size)
                try:
                    t = engine.build_block(text, n, sampling, shape, buf[ref_at:ref_at + int(shape.size)], buf[ssa_at:ssa_at + idx_size])
                except BaseException:
                    bodies.give(buf)
                    raise
                buf[ref_at - len(hb):ref_at] = np.frombuffer(hb, np.uint8)
                buf[ssa_at - len(sb):ssa_at] = np.frombuffer(sb, np.uint8)
                bodies.commit(buf, [(fref.fileno(), ref_off[i], ref_at - len(hb), len(hb) + int(shape.size)),
                                    (fssa.fileno(), ssa_off[i], ssa_at - len(sb), len(sb) + idx_size)])
                t = dict(t or {})
                t["n"], t["block"] = n, i
                timings.append(t)
        finally:
            bodies.close()                                       # the pending writes finish while the files are open
    _barrier(world, group)
    return {"blocks": [[s.header for s in b.sequences] for b in blocks], "owner": owner, "mine": mine,
            "seconds": time.perf_counter() - t0, "timings": timings}


# ---- query -----------------------------------------------------------------------------------------------------------
def _shard_patterns(data, off, lo: int, hi: int):
    """Patterns [lo, hi) of a packed batch, offsets rebased to the shard."""
    off = np.asarray(off)
    b, e = int(off[lo]), int(off[hi])
    sdata = np.ascontiguousarray(np.asarray(data)[b:e]) if e > b else np.zeros(1, np.uint8)
    return sdata, np.ascontiguousarray(off[lo:hi + 1] - b)


def _gather_device(t, *, rank: int, world: int, group=None):
    """One device tensor per rank (equal dtype, any lengths) to rank 0 without leaving the device: lengths first, then one
    padded gather.  Returns the list of tensors on rank 0, None elsewhere."""
    import torch
    dist = _dist()
    mine = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    each = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(world)]
    dist.all_gather(each, mine, group=group)
    lens = [int(x.item()) for x in each]
    width = max(lens)
    if width == 0:
        return [t[:0] for _ in range(world)] if rank == 0 else None
    padded = t.reshape(-1) if t.numel() == width else torch.cat([t.reshape(-1), torch.zeros(width - t.numel(), dtype=t.dtype, device=t.device)])
    parts = [torch.empty(width, dtype=t.dtype, device=t.device) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, parts, dst=0, group=group)
    return [p[:n] for p, n in zip(parts, lens)] if rank == 0 else None


def _is_cuda(device) -> bool:
    return device is not None and str(device).startswith("cuda")


def count_totals_sharded(gssas: Sequence, data, off, *, rank: int = 0, world: int = 1, group=None, device=None):
    """Occurrences of every pattern summed over all blocks (what `gecotools -c PATTERN` logs as "total found",
    tools/GecoMatch.java:110-133), query-sharded: int64[n_patterns] on rank 0, None elsewhere.  On a CUDA `device`
    the shard is uploaded once, searched against every block and summed there (gcz_count_multi), the shards meet on
    rank 0's device through ONE gather, and 8 bytes per pattern come back to the host."""
    n = len(off) - 1
    lo, hi = shard_bounds(n, world)[rank]
    sdata, soff = _shard_patterns(data, off, lo, hi)
    if _is_cuda(device):
        import torch
        from .gssa import count_totals
        tot = torch.zeros(hi - lo, dtype=torch.int64, device=device)
        if hi > lo:
            count_totals(gssas, torch.from_numpy(sdata).to(device), torch.from_numpy(soff).to(device), tot)
        if world == 1:
            return tot.cpu().numpy()
        parts = _gather_device(tot, rank=rank, world=world, group=group)
        return None if parts is None else torch.cat(parts).cpu().numpy()
    local = np.zeros(hi - lo, dtype=np.int64)
    for g in gssas:
        if hi > lo:
            sp, ep = g.count_batch(packed=(sdata, soff))
            local += np.maximum(np.asarray(ep) - np.asarray(sp) + 1, 0)
    parts = gather_varlen(local, rank=rank, world=world, group=group, device=device)
    return None if parts is None else np.concatenate(parts)


def count_sharded(gssas: Sequence, data, off, *, rank: int = 0, world: int = 1, group=None, device=None):
    """Backward-search intervals of a packed pattern batch against every block, query-sharded.

    `gssas`: this rank's replica of the index, one GSSA per block (the same blocks, in the same order, on every
    rank).  Returns on rank 0 two int64 arrays [n_blocks, n_patterns] (sp, ep; ep < sp = not found) — what
    GecoMatch's loop over the blocks (tools/GecoMatch.java:114-131) produces pattern by pattern — and None on
    the other ranks."""
    n = len(off) - 1
    lo, hi = shard_bounds(n, world)[rank]
    sdata, soff = _shard_patterns(data, off, lo, hi)
    k = len(gssas)
    if _is_cuda(device):
        # the shard goes to the device once, is searched there against every block, and the intervals of all ranks meet
        # on rank 0's device before anything is copied to the host
        import torch
        t_out = torch.empty((k, 2, hi - lo), dtype=torch.int64, device=device)
        if hi > lo:
            t_data, t_off = torch.from_numpy(sdata).to(device), torch.from_numpy(soff).to(device)
            for b, g in enumerate(gssas):
                g.count_batch(packed=(t_data, t_off), out=(t_out[b, 0], t_out[b, 1]))
        if world == 1:
            full = t_out.cpu().numpy()
            return full[:, 0, :], full[:, 1, :]
        parts = _gather_device(t_out.reshape(-1), rank=rank, world=world, group=group)
        if parts is None:
            return None
        full = torch.cat([p.reshape(k, 2, -1) for p in parts], dim=2).cpu().numpy()
        return full[:, 0, :], full[:, 1, :]
    local = np.zeros((k, 2, hi - lo), dtype=np.int64)
    for b, g in enumerate(gssas):
        if hi > lo:
            sp, ep = g.count_batch(packed=(sdata, soff))
            local[b, 0], local[b, 1] = np.asarray(sp), np.asarray(ep)
    parts = gather_varlen(local, rank=rank, world=world, group=group, device=device)
    if parts is None:
        return None
    full = np.concatenate([p.reshape(k, 2, -1) for p in parts], axis=2)
    return full[:, 0, :], full[:, 1, :]


def find_sharded(gssas: Sequence, data, off, *, rank: int = 0, world: int = 1, group=None, device=None):
    """GSSA.find (algo/ssa/GSSA.java:160-185) of a packed batch against every block, query-sharded.

    Returns on rank 0 one tuple per block: (per_string_counts [n_patterns, n_strings], positions, pos_off[n+1]) —
    positions pattern after pattern, string after string, ascending and relative to the string start; None on
    the other ranks.  Three gathers per block: counts (fixed size), position totals, then the positions."""
    n = len(off) - 1
    lo, hi = shard_bounds(n, world)[rank]
    sdata, soff = _shard_patterns(data, off, lo, hi)
    out = []
    for g in gssas:
        ns = g.n_strings
        if hi > lo:
            per, pos, poff = g.find_batch_raw(packed=(sdata, soff))
        else:
            per, pos, poff = np.zeros((0, ns), np.int64), np.zeros(0, np.int64), np.zeros(1, np.int64)
        lens = np.diff(np.asarray(poff))
        g_per = gather_varlen(per, rank=rank, world=world, group=group, device=device)
        g_len = gather_varlen(lens, rank=rank, world=world, group=group, device=device)
        g_pos = gather_varlen(pos, rank=rank, world=world, group=group, device=device)
        if rank == 0:
            all_len = np.concatenate(g_len) if g_len else np.zeros(0, np.int64)
            pos_off = np.zeros(n + 1, np.int64)
            np.cumsum(all_len, out=pos_off[1:])
            out.append((np.concatenate([p.reshape(-1, ns) for p in g_per], axis=0) if ns else np.zeros((n, 0), np.int64),
                        np.concatenate(g_pos), pos_off))
    return out if rank == 0 else None
