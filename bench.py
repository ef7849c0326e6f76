#!/usr/bin/env python
"""bench.py — FM-index build Mbp/s (headline) and count queries/s on synthetic ACGTN data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1]

A "step" is one pass of the hot path over one block: gcz_build_block == BlockWriter.run of the reference
(suffix array, BWT, Huffman-shaped wavelet tree, sampled-SA index).  At N=1 the workload is BASELINE.json
configs[1]: a synthetic chr1-shaped block of 248 956 422 bp (+ terminator).  With N>1 (torchrun, one process
per GPU) every rank builds its own chr1-shaped block (seed 3 + rank): blocks are independent, there is no
data-path collective, scaling is weak.

  value  : whole-job Mbp/s with the text and both outputs resident in HBM (device pointers through the C ABI)
  e2e    : same metric through the C-ABI calls of one GecozFileWriter.write (gcz_count_symbols, gcz_shape_from_counts,
           gcz_build_block) with pinned HOST buffers: the H2D of the text and the D2H of the .gcz/.gcx bodies are
           inside the timed region
  count  : backward-search count queries/s against the index just built (secondary metric of BASELINE.json)
  roofline / cpu_baseline : see DESIGN.md
`--impl reference` times the CPU restatement of the Java path (oracle/, no JVM exists on the box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CFG = {
    "cfg2": dict(length=248_956_422, name="cfg2: synthetic chr1-shaped 248956422 bp single block, SA+BWT+HSWT+SSA"),
    "cfg1": dict(length=16_000_000, name="cfg1: synthetic 16 Mbp single sequence"),
}


def make_text(workload: str, rank: int, length: int | None = None) -> np.ndarray:
    from gecoz_b200 import synth
    ln = length or CFG[workload]["length"]
    if workload == "cfg1":
        return synth.block_of([synth.iid_acgtn(ln, seed=1 + rank)])
    return synth.cfg2_text(ln, seed=3 + rank)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 50 ms from before the warm-up on; stop(t0, t1) keeps the
    samples taken inside the timed region [t0, t1] (wall clock), or the nearest ones when the region is shorter
    than a sampling period."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.rows = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.06]
        how = "inside the timed region"
        if not rows and self.rows:
            mid = 0.5 * (t0 + t1)
            rows = [r for _, r in sorted(self.rows, key=lambda tr: abs(tr[0] - mid))[:3]]
            how = "nearest to the timed region (region shorter than the sampling period)"
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "sampled": how}


# ---------------------------------------------------------------------------------------------------------------
def run_reference(args, rank: int) -> None:
    """CPU arm: the oracle's restatement of BlockWriter.run (SA-IS, then HSWT on a side thread while the
    sampled-SA index is written, as fmt/GecozFileWriter.java:264-277) on a bounded sample."""
    if rank != 0:
        return
    from oracle import gcz_oracle as O
    O.build()
    steps, warm = max(1, args.steps), max(0, args.warmup)
    budget_s = 150.0 / (steps + warm)
    sample = int(min(64_000_000, max(2_000_000, budget_s * 4.0e6)))
    text = make_text(args.workload, 0, sample)
    bases = len(text) - 1
    times = []
    for i in range(warm + steps):
        t0 = time.perf_counter()
        O.build_block(text, 32, threads=2)
        if i >= warm:
            times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    value = bases / 1e6 / sec
    # supplementary: the host saturated the way `gecotools -t N` saturates it on a multi-block genome — cores / 2 blocks in
    # flight, two threads each (the config of this arm is ONE block, which the reference cannot spread further)
    from concurrent.futures import ThreadPoolExecutor
    conc = max(1, (os.cpu_count() or 2) // 2)
    small = make_text(args.workload, 0, max(2_000_000, sample // 4))
    t0 = time.perf_counter()
    with ThreadPoolExecutor(conc) as pool:
        list(pool.map(lambda _: O.build_block(small, 32, threads=2), range(conc)))
    all_cores = {"blocks_in_flight": conc, "threads": 2 * conc, "block_bp": len(small) - 1,
                 "value": conc * (len(small) - 1) / 1e6 / (time.perf_counter() - t0), "unit": "Mbp/s"}
    line = {
        "impl": "reference", "metric": "FM-index build throughput (SA+BWT+HSWT+SSA per block)", "value": value, "unit": "Mbp/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
        "config": {"workload": CFG[args.workload]["name"], "sample": f"{sample} bp block of the same generator (scaled N runs)"},
        "cpu_baseline": {"value": value, "unit": "Mbp/s", "cores": 2, "kind": "port",
                         "sample": f"{sample} bp chr1-shaped block; C restatement of the Java path (no JVM on this box), "
                                   f"SA-IS single-threaded then HSWT || SSA on 2 threads like BlockWriter.run; host has {os.cpu_count()} cores, "
                                   f"one block can use 2", "all_cores_multi_block": all_cores},
        "e2e": {"value": value, "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(CFG))
    ap.add_argument("--length", type=int, default=None, help="override the block length (debugging)")
    ap.add_argument("--patterns", type=int, default=4_000_000, help="count-leg patterns (length 15..100)")
    ap.add_argument("--locate-patterns", type=int, default=200_000, help="locate-leg patterns (subset of the count batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import gecoz_b200 as G
    from gecoz_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: gecoz_b200 has no CPU fallback")
    G.lib()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(minutes=10))   # > the 300 s deadline below
    steps, warm = max(1, args.steps), max(3, args.warmup)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- workload -------------------------------------------------------------------------------------------
    text = make_text(args.workload, rank, args.length)
    n = len(text)
    bases = n - 1
    # a dedicated (non-default) stream: the library launches on it, and the timing events are recorded on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    G._native.check(G.lib().gcz_set_stream(local_rank, stream.cuda_stream))
    h_text = torch.from_numpy(text).pin_memory()
    d_text = h_text.to(dev, non_blocking=True)
    shape = G.shape_from_counts(G.symbol_counts(d_text, local_rank))
    gcx_len = G.index_size(n, 5)
    d_gcz = torch.empty(int(shape.size), dtype=torch.uint8, device=dev)
    d_gcx = torch.empty(gcx_len, dtype=torch.uint8, device=dev)
    h_gcz = torch.empty(int(shape.size), dtype=torch.uint8).pin_memory()
    h_gcx = torch.empty(gcx_len, dtype=torch.uint8).pin_memory()

    def timed(fn, k: int) -> tuple[float, list[dict]]:
        """k steps bracketed by barrier + synchronize, device time from CUDA events on the launching stream."""
        infos = []
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k):
            infos.append(fn())
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1), infos

    dev_step = lambda: G.build_block(local_rank, d_text, n, 32, shape, d_gcz, d_gcx)

    def stage():
        # the per-block head of GecozFileWriter.write from host memory: count (= the upload, kept on the device), shape
        return G.shape_from_counts(G.symbol_counts(h_text, local_rank))

    def e2e_step():
        return G.build_block(local_rank, h_text, n, 32, stage(), h_gcz, h_gcx)

    from concurrent.futures import ThreadPoolExecutor
    stager = ThreadPoolExecutor(1)

    def e2e_pipelined(k: int):
        # what GecozFileWriter does with its two blocks in flight per GPU: block i + 1 is counted / uploaded (the
        # library's staging stream) while block i is being built; every step still moves its own text in and its
        # own bodies out, inside the timed region
        nxt = stager.submit(stage)
        for i in range(k):
            shp = nxt.result()
            if i + 1 < k:
                nxt = stager.submit(stage)
            G.build_block(local_rank, h_text, n, 32, shp, h_gcz, h_gcx)

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    timed(dev_step, warm)
    w0 = time.time()
    ms_total, infos = timed(dev_step, steps)
    clk = clocks.stop(w0, time.time()) if rank == 0 else None
    ms_step = max_over_ranks(ms_total / steps)
    total_bases = sum_over_ranks(float(bases))
    value = total_bases / 1e6 / (ms_step / 1e3)

    timed(e2e_step, 1)
    ms_serial_total, _ = timed(e2e_step, steps)
    ms_e2e_serial = max_over_ranks(ms_serial_total / steps)
    timed(lambda: e2e_pipelined(2), 1)
    t0 = time.perf_counter()
    ms_e2e_total, _ = timed(lambda: e2e_pipelined(steps), 1)
    wall_e2e_ms = (time.perf_counter() - t0) * 1e3
    # device events on the build stream do not see a staging that runs ahead of the first build: take the longer of the two clocks
    ms_e2e = max_over_ranks(max(ms_e2e_total, wall_e2e_ms if world == 1 else ms_e2e_total) / steps)
    assert torch.equal(h_gcz, d_gcz.cpu()) and torch.equal(h_gcx, d_gcx.cpu()), "device and host arms disagree"
    e2e_value = total_bases / 1e6 / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel: an onesweep digit pass over all n (key, value) pairs ---------------------------
    # achieved = algorithmic bytes of one such launch (24 B per pair: 12 read + 12 written) / its average device time,
    # measured live with CUDA events around every launch inside the library (on the stream it launches on).
    peak, peak_src = measured_peaks()
    full_ms = float(np.mean([i["radix_full_ms"] for i in infos]))
    full_launches = float(np.mean([i["radix_full_launches"] for i in infos]))
    all_ms = float(np.mean([i["radix_ms"] for i in infos]))
    all_launches = float(np.mean([i["radix_launches"] for i in infos]))
    text_ms = float(np.mean([i["radix_text_ms"] for i in infos]))
    alg_bytes_per_launch = 24.0 * n
    avg_launch_ms = full_ms / max(full_launches, 1)
    achieved = alg_bytes_per_launch / (avg_launch_ms / 1e3) / 1e9 if avg_launch_ms > 0 else 0.0
    step_alg_bytes = 11.0 * n + int(shape.size) + gcx_len                          # SURVEY.md §8(d) B_build(n)
    roofline = {
        "bound": "hbm", "kernel": "onesweep_kernel<512,12,pairs>: one 8-bit digit pass of the suffix sorter over all n (key, position) pairs",
        "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
        "traffic": None,
        "algorithmic_bytes_per_launch": alg_bytes_per_launch, "launches_per_step": full_launches,
        "avg_launch_ms": avg_launch_ms, "kernel_share_of_step": full_ms / (ms_total / steps),
        "all_digit_passes": {"launches_per_step": all_launches, "ms_per_step": all_ms, "share_of_step": all_ms / (ms_total / steps),
                             "first_pass_from_text_ms": text_ms},
        "whole_step": {"algorithmic_bytes": step_alg_bytes, "achieved": step_alg_bytes / (ms_total / steps / 1e3) / 1e9,
                       "frac": step_alg_bytes / (ms_total / steps / 1e3) / 1e9 / peak},
    }
    tr = ROOT / "profiles" / "traffic_r01.json"
    if tr.exists():
        try:
            roofline["traffic"] = json.loads(tr.read_text()).get("onesweep_bytes_per_launch")
        except Exception:
            pass

    line = None
    if rank == 0:
        line = {
            "metric": "FM-index build throughput (SA+BWT+HSWT+SSA per block)", "value": value, "unit": "Mbp/s",
            "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32 (64-bit packed keys)", "data": "synthetic",
            "config": {"workload": CFG[args.workload]["name"] + (f" (length overridden to {args.length})" if args.length else ""),
                       "symbols_per_block": n, "blocks": world, "sampling_rate": 32,
                       "l2": "inputs larger than L2 (249 MB text, ~3 GB sort working set per pass vs 126 MB L2)",
                       "parallelism": f"{world} independent block(s), one per GPU, no collective"},
            "e2e": {"value": e2e_value, "unit": "Mbp/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(n),
                    "d2h_bytes_per_step": int(shape.size) + gcx_len,
                    "pipelining": "the upload + histogram of step i + 1 overlaps the build of step i (two text slots per device), as in "
                                  "GecozFileWriter; K steps timed as one region",
                    "serial": {"value": total_bases / 1e6 / (ms_e2e_serial / 1e3), "ms_per_step": ms_e2e_serial,
                               "what": "the same calls strictly one after the other"}},
            "gpu_launches": int(sum(i["kernel_launches"] for i in infos)),
            "clocks": clk,
            "roofline": roofline,
            "cpu_baseline": None, "count": None, "locate": None,
            "phases_ms": {k: float(np.mean([i[k] for i in infos])) for k in
                          ("sort_initial_ms", "sort_refine_ms", "bwt_hswt_ms", "ssa_ms", "total_ms")},
            "refine_rounds": int(infos[-1]["refine_rounds"]),
            "sorter": {"symbols_per_key": int(infos[-1]["symbols_per_key"]), "long_runs": int(infos[-1]["long_runs"]),
                       "unresolved_after_first_sort": int(infos[-1]["unresolved_after_first_sort"])},
        }

    # The headline numbers are complete here.  The query legs below use collectives; if one of them hangs (a rank that
    # died, a lost peer) the line is still printed and every rank leaves with status 0 instead of waiting for NCCL's
    # watchdog to abort the job.
    def give_up():
        if rank == 0:
            line["count"] = line["count"] or {"error": "query legs did not finish within the deadline"}
            print(json.dumps(line), flush=True)
        os._exit(0)

    deadline = threading.Timer(300.0, give_up)
    deadline.daemon = True
    deadline.start()

    # ---- count leg: query-sharded batch against a replicated index ---------------------------------------
    # The index of rank 0's block is replicated (NCCL broadcast of the two bodies); the global batch of
    # world x --patterns patterns is cut into contiguous shards; every rank counts its shard and the intervals
    # come back to rank 0 with one NCCL gather inside the timed region.
    count = locate = None
    try:
        from gecoz_b200 import sharding
        if world > 1:
            meta = torch.tensor([int(shape.size), gcx_len, n], dtype=torch.int64, device=dev)
            dist.broadcast(meta, 0)
            size0, gcx0, n0 = (int(x) for x in meta.tolist())
            r_gcz = d_gcz if rank == 0 else torch.empty(size0, dtype=torch.uint8, device=dev)
            r_gcx = d_gcx if rank == 0 else torch.empty(gcx0, dtype=torch.uint8, device=dev)
            dist.broadcast(r_gcz, 0)
            dist.broadcast(r_gcx, 0)
        else:
            r_gcz, r_gcx, n0 = d_gcz, d_gcx, n
        g = G.GSSA.open(local_rank, r_gcz, n0, r_gcx)
        per_rank = max(1000, args.patterns)
        npat = per_rank * world
        lo, hi = sharding.shard_bounds(npat, world)[rank]
        width = per_rank
        # shard r of the batch is drawn by rank r itself (seed 5 + r) from the text of the replicated block, which every
        # rank regenerates from its seed: no rank materialises the whole batch and nothing is sent
        text0 = text if rank == 0 else make_text(args.workload, 0, args.length)
        sdata, soff = synth.patterns(text0, per_rank, 15, 100, seed=5 + rank)
        # rank 0 also draws shard 1: what it computes for it is compared with rank 1's gathered result below
        keep = synth.patterns(text0, per_rank, 15, 100, seed=5 + 1) if (rank == 0 and world > 1) else None
        del text0
        hp, ho = torch.from_numpy(sdata).pin_memory(), torch.from_numpy(soff).pin_memory()
        dp, do = hp.to(dev), ho.to(dev)
        d_res = torch.full((2, width), -1, dtype=torch.int64, device=dev)        # row 0 = sp, row 1 = ep
        h_res = torch.empty((2, width), dtype=torch.int64).pin_memory()
        parts = [torch.empty((2, width), dtype=torch.int64, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None
        h_all = torch.empty((world, 2, width), dtype=torch.int64).pin_memory() if rank == 0 else None

        def cdev():
            g.count_batch(packed=(dp, do), out=(d_res[0, :hi - lo], d_res[1, :hi - lo]))
            if world > 1:
                dist.gather(d_res, parts, dst=0)

        def chost():
            # host shard in, all intervals back on rank 0's host
            g.count_batch(packed=(hp, ho), out=(h_res[0, :hi - lo], h_res[1, :hi - lo]))
            if world > 1:
                d_res.copy_(h_res, non_blocking=True)
                dist.gather(d_res, parts, dst=0)
                if rank == 0:
                    h_all.copy_(torch.stack(parts), non_blocking=True)

        timed(cdev, 3)
        cms, _ = timed(cdev, 5)
        cms = max_over_ranks(cms / 5)
        timed(chost, 1)
        cms_e2e, _ = timed(chost, 5)
        cms_e2e = max_over_ranks(cms_e2e / 5)
        torch.cuda.synchronize()
        assert torch.equal(h_res[:, :hi - lo], d_res[:, :hi - lo].cpu())
        if rank == 0 and world > 1:                        # a gathered shard == what rank 0 computes for it itself
            chk_sp, chk_ep = g.count_batch(packed=keep)
            got = parts[1].cpu().numpy()
            assert np.array_equal(got[0], chk_sp) and np.array_equal(got[1], chk_ep), "sharded count differs"
        found = int(sum_over_ranks(float((d_res[1, :hi - lo] >= d_res[0, :hi - lo]).sum().item())))
        h2d_patterns = int(sum_over_ranks(float(sdata.nbytes + soff.nbytes)))
        count = {"metric": "count queries/s (backward-search intervals)", "value": npat / (cms / 1e3), "unit": "queries/s",
                 "patterns": int(npat), "pattern_length": "uniform 15..100, 50% text-sampled / 50% random", "found": found,
                 "ms_per_batch": cms, "sharding": f"{world} contiguous shard(s), replicated index" + (", one NCCL gather to rank 0" if world > 1 else ""),
                 "e2e": {"value": npat / (cms_e2e / 1e3), "unit": "queries/s", "h2d_bytes_per_step": h2d_patterns,
                         "d2h_bytes_per_step": int(npat * 16)}}

        # ---- locate leg (rank 0's shard only at N>1 is not the point: every rank runs its shard, no gather timed) ----
        nloc = max(1000, min(args.locate_patterns, hi - lo))
        ldata, loff = sharding._shard_patterns(sdata, soff, 0, nloc)
        t0 = time.perf_counter()
        per, pos, pof = g.find_batch_raw(packed=(ldata, loff))
        torch.cuda.synchronize()
        lsec = time.perf_counter() - t0
        t0 = time.perf_counter()
        per, pos, pof = g.find_batch_raw(packed=(ldata, loff))
        lsec = min(lsec, time.perf_counter() - t0)
        locate = {"metric": "locate (GSSA.find) through gcz_find_batch, host buffers, wall clock of the call", "patterns": int(nloc),
                  "occurrences": int(len(pos)), "patterns_per_s": nloc / lsec, "occurrences_per_s": len(pos) / lsec, "ms": lsec * 1e3,
                  "per_rank": True}
        g.close()
    except Exception as ex:                                  # the headline metric must still be reported
        count = count or {"error": repr(ex)}
        locate = locate or {"error": repr(ex)}

    deadline.cancel()

    # ---- CPU baseline (rank 0, N=1 only) --------------------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import gcz_oracle as O
        sample = min(64_000_000, bases)
        stext = make_text(args.workload, 0, sample) if sample < bases else text
        t0 = time.perf_counter()
        O.build_block(stext, 32, threads=2)
        sec = time.perf_counter() - t0
        cpu = {"value": (len(stext) - 1) / 1e6 / sec, "unit": "Mbp/s", "cores": 2, "kind": "port",
               "sample": f"{len(stext) - 1} bp chr1-shaped block from the same generator, one run ({sec:.1f} s); C restatement of the "
                         f"Java path (no JVM on the box); one block can use 2 threads (SA-IS, then HSWT || SSA); host has {os.cpu_count()} cores"}

    if rank == 0:
        line["cpu_baseline"], line["count"], line["locate"] = cpu, count, locate
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
