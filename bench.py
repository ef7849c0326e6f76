#!/usr/bin/env python
"""bench.py — FM-index build Mbp/s (headline) and count queries/s on synthetic ACGTN data of the BASELINE.json shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input:

  N = 1   headline = BASELINE.json configs[1] (cfg2): one gcz_build_block == BlockWriter.run of the reference (suffix array, BWT,
          Huffman-shaped wavelet tree, sampled-SA index) on a chr1-shaped block of 248 956 422 bp (+ terminator).  The same line
          carries the N = 1 point of the multi-GPU curves: `genome` (cfg3, all 18 blocks on this GPU), `count` (cfg4), `locate` (cfg5).
  N > 1   headline = configs[2] (cfg3): the 25 hg38-length sequences -> the reference's 18 chromosome-bounded blocks
          (tools/GecoIndex.java:72-98), block -> rank by longest-processing-time-first, no data-path collective ("scaling":
          "strong": the genome is fixed, a step = every rank builds its share once).  Then configs[3]: the index is replicated
          (NCCL broadcast of every block's bodies from its builder), the pattern batch is cut into one shard per rank, every rank
          counts its shard against all 18 blocks, and the per-pattern totals come back to rank 0 with ONE NCCL gather inside the
          timed region; configs[4]: GSSA.find of a sharded batch, occurrences located and split per string.

  value     whole-job Mbp/s with the text and both outputs resident in HBM (device pointers through the C ABI)
  e2e       the same metric through the C-ABI calls of GecozFileWriter.write (gcz_count_symbols, gcz_shape_from_counts,
            gcz_build_block) with pinned HOST buffers: H2D of the text and D2H of the .gcz/.gcx bodies inside the timed region
  roofline  the dominant kernel (a full-size digit pass of the suffix sorter) against the measured HBM peak; `count.roofline` the
            backward-search kernel (32 B rank sectors actually read; the reference layout's 74 B per rank call beside it)
  parity    sha256 of every block body built in the e2e pass == the oracle's digest (tests/golden/full_size_digests.json,
            made by tools/make_digests.py): byte parity at the benchmarked size, checked inside the bench run

`--impl reference` times the CPU restatement of the Java path (oracle/; no JVM exists on the box) on the SAME config: the full
cfg2 block at N = 1 (SA-IS, then HSWT || SSA on two threads like BlockWriter.run), the whole hg38-shaped genome with every host
core at N > 1 (rank 0 only).  Any failure on any rank prints that rank's traceback to stderr and ends the job with a non-zero
status; a phase that outlives its deadline dumps every thread's stack and exits 3.
"""
from __future__ import annotations

import argparse
import faulthandler
import hashlib
import json
import os
import subprocess
import sys
import threading
import time
import traceback
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CFG2_LEN = 248_956_422
METRIC = "FM-index build throughput (SA+BWT+HSWT+SSA per block)"
GOLDEN = ROOT / "tests" / "golden" / "full_size_digests.json"


def log(rank: int, msg: str) -> None:
    print(f"[bench rank {rank} +{time.time() - T_START:6.1f}s] {msg}", file=sys.stderr, flush=True)


T_START = time.time()


class Watchdog:
    """Every phase gets a deadline.  A phase that outlives it is a hang (a lost peer in a collective, a wedged kernel): all
    thread stacks go to stderr and the process exits 3, which makes torchrun tear the job down with a non-zero status."""

    def __init__(self, rank: int):
        self.rank, self.timer, self.phase = rank, None, "start"

    def enter(self, phase: str, seconds: float) -> None:
        self.cancel()
        self.phase = phase
        self.timer = threading.Timer(seconds, self._fire, args=(phase, seconds))
        self.timer.daemon = True
        self.timer.start()
        log(self.rank, f"phase: {phase}")

    def cancel(self) -> None:
        if self.timer is not None:
            self.timer.cancel()
            self.timer = None

    def _fire(self, phase: str, seconds: float) -> None:
        print(f"[bench rank {self.rank}] DEADLINE: phase '{phase}' did not finish within {seconds:.0f} s; thread stacks follow",
              file=sys.stderr, flush=True)
        faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        sys.stderr.flush()
        os._exit(3)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def host_info() -> dict:
    info = {"cores": os.cpu_count()}
    try:
        import psutil
        vm = psutil.virtual_memory()
        info["ram_gb"] = round(vm.total / 2 ** 30, 1)
        info["ram_available_gb"] = round(vm.available / 2 ** 30, 1)
    except Exception:
        pass
    return info


def sha(a) -> str:
    return hashlib.sha256(memoryview(np.ascontiguousarray(a)).cast("B")).hexdigest()


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


# ---- workloads ---------------------------------------------------------------------------------------------------------------
def genome_plan(scale: float):
    """The hg38-shaped genome as the reference would block it: [(headers, member ids, symbols)] in file order
    (tools/GecoIndex.java:72-98 through the product's own merge)."""
    from gecoz_b200 import synth
    from gecoz_b200.geco_index import FastaSequence, merge_blocks
    lengths = [max(8, int(ln * scale)) for ln in synth.HG38_LENGTHS]
    seqs = [FastaSequence(h, ln, None, i) for i, (h, ln) in enumerate(zip(synth.HG38_NAMES, lengths))]
    return lengths, [([s.header for s in b.sequences], [s.id for s in b.sequences], int(b.size)) for b in merge_blocks(seqs)]


def block_text_into(dst: np.ndarray, lengths, ids) -> None:
    """Members in block order, each followed by '\\0' (tools/GecoIndex.java:138-142), synthesised straight into `dst`."""
    from gecoz_b200 import synth
    p = 0
    for i in ids:
        dst[p:p + lengths[i]] = synth.chromosome_shaped(lengths[i], 4 + i)
        dst[p + lengths[i]] = 0
        p += lengths[i] + 1
    assert p == len(dst)


def config_of(world: int, scale: float, steps_note: str | None = None) -> dict:
    """The `config` object — identical in both arms (ours / reference) for the same N."""
    if world == 1:
        return {"workload": f"cfg2: synthetic chr1-shaped {CFG2_LEN} bp single block, SA+BWT+HSWT+SSA",
                "symbols_per_block": CFG2_LEN + 1, "blocks": 1, "sampling_rate": 32,
                "l2": "inputs larger than L2 (249 MB text, ~3 GB sort working set per pass vs 126 MB L2)",
                "parallelism": "1 block on 1 GPU"}
    lengths, plan = genome_plan(scale)
    return {"workload": f"cfg3: synthetic hg38-shaped genome, 25 sequences, {sum(lengths)} bp -> {len(plan)} chromosome-bounded blocks"
                        + (f" (lengths scaled by {scale})" if scale != 1.0 else ""),
            "symbols": int(sum(lengths) + len(lengths)), "blocks": len(plan), "sampling_rate": 32,
            "l2": "inputs larger than L2 (blocks of 116-249 M symbols, ~1.4-3 GB sort working set per pass vs 126 MB L2)",
            "parallelism": f"blocks -> {world} ranks by LPT, one process per GPU, no data-path collective in the build"}


# ---- reference arm ---------------------------------------------------------------------------------------------------------------
def run_reference(args, rank: int, world: int) -> None:
    """The CPU arm: the oracle's restatement of the Java path on the SAME config as our arm.
    N = 1: BlockWriter.run on the full cfg2 block (SA-IS single-threaded, then HSWT || SSA on 2 threads,
    fmt/GecozFileWriter.java:256-284 — one block cannot use more).  N > 1: GecoIndex.index over the whole hg38-shaped
    genome with `-t cores`: WriterPoolExecutor keeps cores / 2 blocks in flight, two threads each (:174-227)."""
    if rank != 0:
        return
    from gecoz_b200 import synth
    from oracle import gcz_oracle as O
    O.build()
    steps, warm = max(1, args.steps), max(0, args.warmup)
    budget = float(args.ref_budget)
    cores = os.cpu_count() or 2
    t_begin = time.time()
    if world == 1:
        text = synth.cfg2_text(args.length or CFG2_LEN, seed=3)
        bases = len(text) - 1
        threads_used = 2

        def one_step():
            O.build_block(text, 32, threads=2)
        sample = f"the full {bases} bp chr1-shaped block (same generator and seed as the GPU arm)"
    else:
        lengths, plan = genome_plan(args.scale)
        texts = []
        with ThreadPoolExecutor(min(8, cores)) as pool:
            def make(b):
                t = np.empty(plan[b][2], np.uint8)
                block_text_into(t, lengths, plan[b][1])
                return t
            texts = list(pool.map(make, range(len(plan))))
        bases = int(sum(lengths))
        in_flight = max(1, min(len(plan), cores // 2))
        threads_used = 2 * in_flight

        def one_step():
            order = sorted(range(len(plan)), key=lambda b: -plan[b][2])
            with ThreadPoolExecutor(in_flight) as pool:
                list(pool.map(lambda b: O.build_block(texts[b], 32, threads=2), order))
        sample = (f"the whole hg38-shaped genome ({bases} bp, {len(plan)} blocks), {in_flight} blocks in flight x 2 threads "
                  f"= what `gecotools -t {cores}` keeps busy")
    # the CPU needs no warm-up to reach steady state: at most one untimed step, and the timed steps stop at the time budget
    times = []
    for i in range(min(warm, 1) + steps):
        t0 = time.perf_counter()
        one_step()
        dt = time.perf_counter() - t0
        if i >= min(warm, 1):
            times.append(dt)
        if time.time() - t_begin + dt > budget and times:
            break
    sec = float(np.mean(times))
    value = bases / 1e6 / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mbp/s",
        "n_gpus": args.gpus, "steps": len(times), "steps_requested": steps, "warmup": min(warm, 1), "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
        "config": config_of(world, args.scale),
        "cpu_baseline": {"value": value, "unit": "Mbp/s", "cores": threads_used, "kind": "port",
                         "sample": sample + f"; C restatement of the Java path (no JVM on this box); host: {host_info()}; "
                                            f"{len(times)} timed step(s) of {steps} requested (time budget {budget:.0f} s)"},
        "e2e": {"value": value, "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---- our arm ---------------------------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import gecoz_b200 as G
        self.torch, self.dist, self.G, self.args = torch, dist, G, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.wd = Watchdog(self.rank)
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: gecoz_b200 has no CPU fallback")
        G.lib()
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            import datetime
            self.wd.enter("init_process_group", 240)
            dist.init_process_group("nccl", device_id=self.dev, timeout=datetime.timedelta(minutes=8))
        self.steps, self.warm = max(1, args.steps), max(3, args.warmup)
        # a dedicated (non-default) stream: the library launches on it, and the timing events are recorded on it
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        G._native.check(G.lib().gcz_set_stream(self.local_rank, self.stream.cuda_stream))
        self.peak, self.peak_src = measured_peaks()
        self.gold = json.loads(GOLDEN.read_text()) if GOLDEN.exists() else {}
        # the writer's pool: two blocks in flight per device (fmt/GecozFileWriter.java:174-227); its threads launch on the timed stream
        self.builders = ThreadPoolExecutor(2, initializer=lambda: G._native.check(G.lib().gcz_set_stream(self.local_rank, self.stream.cuda_stream)))
        log(self.rank, f"world {self.world}, device {self.local_rank}, host {host_info()}")

    # -- plumbing ----------------------------------------------------------------------------------------------------------------
    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, x: float, op: str) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN}[op])
        return float(t.item())

    def timed(self, fn, k: int):
        """k steps bracketed by barrier + synchronize on both sides, device time from CUDA events on the launching stream."""
        torch = self.torch
        out = []
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(k):
            out.append(fn())
        e1.record(self.stream)
        self.barrier()
        return e0.elapsed_time(e1), out

    def pinned(self, nbytes: int):
        return self.torch.empty(int(nbytes), dtype=self.torch.uint8).pin_memory()

    def roofline_of(self, infos, n_for_launch: int, step_ms: float, step_alg_bytes: float, nsteps: int) -> dict:
        """Dominant kernel = a digit pass of the suffix sorter over all n (key, position) pairs of a block: 24 B per pair
        (12 read + 12 written).  Its device time is measured live (CUDA events around every digit pass inside the
        library, on the stream it launches on)."""
        full_ms = float(np.sum([i["radix_full_ms"] for i in infos]))
        full_launches = float(np.sum([i["radix_full_launches"] for i in infos]))
        full_elems = float(np.sum([i["radix_full_launches"] * i["_n"] for i in infos]))
        all_ms = float(np.sum([i["radix_ms"] for i in infos]))
        alg_bytes = 24.0 * full_elems                              # over every full-size launch of the timed region
        achieved = alg_bytes / (full_ms / 1e3) / 1e9 if full_ms > 0 else 0.0
        r = {"bound": "hbm",
             "kernel": "onesweep digit pass of the suffix sorter over all n (key, position) pairs of a block",
             "achieved": achieved, "peak": self.peak, "peak_source": self.peak_src, "unit": "GB/s", "frac": achieved / self.peak,
             "traffic": None,
             "algorithmic_bytes_per_launch": 24.0 * n_for_launch, "launches_in_timed_region": full_launches,
             "avg_launch_ms": full_ms / max(full_launches, 1) * (24.0 * n_for_launch) / max(alg_bytes / max(full_launches, 1), 1),
             "kernel_share_of_step": full_ms / max(step_ms * nsteps, 1e-9),
             "all_digit_passes": {"ms": all_ms, "share_of_step": all_ms / max(step_ms * nsteps, 1e-9)},
             "whole_step": {"algorithmic_bytes": step_alg_bytes, "achieved": step_alg_bytes / (step_ms / 1e3) / 1e9,
                            "frac": step_alg_bytes / (step_ms / 1e3) / 1e9 / self.peak}}
        for name in ("traffic_r02.json", "traffic_r01.json"):
            tr = ROOT / "profiles" / name
            if tr.exists():
                try:
                    d = json.loads(tr.read_text())
                    r["traffic"] = d.get("onesweep_bytes_per_launch")
                    r["traffic_source"] = f"profiles/{name}: dram bytes read + written of one launch over {d.get('pairs', CFG2_LEN + 1)} pairs (ncu --set full)"
                    break
                except Exception:
                    pass
        return r

    # -- leg 1 (N = 1): the cfg2 block ------------------------------------------------------------------------------------------
    def leg_block(self) -> dict:
        torch, G = self.torch, self.G
        from gecoz_b200 import synth
        dev, lr = self.dev, self.local_rank
        self.wd.enter("cfg2: synthesis + upload", 300)
        text = synth.cfg2_text(self.args.length or CFG2_LEN, seed=3)
        n = len(text)
        bases = n - 1
        h_text = torch.from_numpy(text).pin_memory()
        d_text = h_text.to(dev, non_blocking=True)
        shape = G.shape_from_counts(G.symbol_counts(d_text, lr))
        gcx_len = G.index_size(n, 5)
        d_gcz = torch.empty(int(shape.size), dtype=torch.uint8, device=dev)
        d_gcx = torch.empty(gcx_len, dtype=torch.uint8, device=dev)
        h_gcz, h_gcx = self.pinned(int(shape.size)), self.pinned(gcx_len)
        # the writer keeps two blocks in flight: two sets of pinned host buffers (same text)
        w_text = [h_text, h_text.clone().pin_memory()]
        w_gcz, w_gcx = [h_gcz, self.pinned(int(shape.size))], [h_gcx, self.pinned(gcx_len)]

        def dev_step():
            t = G.build_block(lr, d_text, n, 32, shape, d_gcz, d_gcx)
            t["_n"] = n
            return t

        def stage(w=0):
            # the per-block head of GecozFileWriter.write from host memory: count (= the upload, kept on the device), shape
            return G.shape_from_counts(G.symbol_counts(w_text[w], lr))

        def e2e_step():
            return G.build_block(lr, h_text, n, 32, stage(), h_gcz, h_gcx)

        e2e_infos = []

        def e2e_writer(k: int):
            # GecozFileWriter.write + its pool (fmt/GecozFileWriter.java:124-159, 174-227): the submitting thread counts block i
            # (= its upload; the text stays staged on the device), then queues it; two pool threads run BlockWriter.run, which
            # the device serialises — block i + 1 is uploaded while block i is built, and the bodies of block i travel to the
            # host while block i + 1 is built.  Every step moves its own text in and its own bodies out, inside the timed region.
            pending = []
            for i in range(k):
                w = i % 2
                while len(pending) >= 2:
                    e2e_infos.append(pending.pop(0).result())
                shp = stage(w)
                pending.append(self.builders.submit(G.build_block, lr, w_text[w], n, 32, shp, w_gcz[w], w_gcx[w]))
            for f in pending:
                e2e_infos.append(f.result())

        self.wd.enter("cfg2: device-resident steps", 300)
        clocks = ClockSampler(lr)
        clocks.start()
        self.timed(dev_step, self.warm)
        w0 = time.time()
        ms_total, infos = self.timed(dev_step, self.steps)
        clk = clocks.stop(w0, time.time())
        ms_step = ms_total / self.steps
        value = bases / 1e6 / (ms_step / 1e3)

        self.wd.enter("cfg2: e2e steps", 300)
        self.timed(e2e_step, 1)
        ms_serial_total, _ = self.timed(e2e_step, self.steps)
        self.timed(lambda: e2e_writer(3), 1)
        del e2e_infos[:]
        t0 = time.perf_counter()
        ms_e2e_total, _ = self.timed(lambda: e2e_writer(self.steps), 1)
        wall_e2e_ms = (time.perf_counter() - t0) * 1e3
        # device events on the build stream do not see a staging that runs ahead of the first build: take the longer of the two clocks
        ms_e2e = max(ms_e2e_total, wall_e2e_ms) / self.steps
        for w in (0, 1):
            assert torch.equal(w_gcz[w], d_gcz.cpu()) and torch.equal(w_gcx[w], d_gcx.cpu()), "device and host arms disagree"

        # byte parity at the benchmarked size: the bodies the e2e arm just wrote against the oracle's digests
        parity = {"checked": False, "why": "no golden digests for this length"}
        g2 = self.gold.get("cfg2")
        if g2 and g2["n"] == n:
            got = {"gcz_body": sha(h_gcz.numpy()), "gcx_body": sha(h_gcx.numpy())}
            ok = got["gcz_body"] == g2["gcz_body"] and got["gcx_body"] == g2["gcx_body"]
            parity = {"checked": True, "ok": ok, "against": "oracle digests, tests/golden/full_size_digests.json (tools/make_digests.py)", **got}
            if not ok:
                raise AssertionError(f"cfg2 bodies differ from the oracle's digests: {got} vs {g2['gcz_body']}, {g2['gcx_body']}")

        step_alg_bytes = 11.0 * n + int(shape.size) + gcx_len                          # SURVEY.md §8(d) B_build(n)
        out = {
            "value": value, "ms_per_step": ms_step, "clocks": clk,
            "e2e": {"value": bases / 1e6 / (ms_e2e / 1e3), "unit": "Mbp/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(n),
                    "d2h_bytes_per_step": int(shape.size) + gcx_len,
                    "pipelining": "GecozFileWriter's schedule: two blocks in flight per device — the upload + histogram of step i + 1 and the "
                                  "copies of step i - 1's bodies to the host overlap the kernels of step i; K steps timed as one region",
                    "serial": {"value": bases / 1e6 / (ms_serial_total / self.steps / 1e3), "ms_per_step": ms_serial_total / self.steps,
                               "what": "the same calls strictly one after the other"},
                    "build_call_phases_ms": {k: float(np.mean([i[k] for i in e2e_infos[-self.steps:]])) for k in
                                             ("h2d_ms", "sort_initial_ms", "sort_refine_ms", "bwt_hswt_ms", "ssa_ms", "d2h_ms", "total_ms")}},
            "gpu_launches": int(sum(i["kernel_launches"] for i in infos)),
            "roofline": self.roofline_of(infos, n, ms_step, step_alg_bytes, self.steps),
            "parity": parity,
            "phases_ms": {k: float(np.mean([i[k] for i in infos])) for k in ("sort_initial_ms", "sort_refine_ms", "bwt_hswt_ms", "ssa_ms", "total_ms")},
            "refine_rounds": int(infos[-1]["refine_rounds"]),
            "sorter": {"symbols_per_key": int(infos[-1]["symbols_per_key"]), "long_runs": int(infos[-1]["long_runs"]),
                       "unresolved_after_first_sort": int(infos[-1]["unresolved_after_first_sort"])},
        }
        del d_text, d_gcz, d_gcx, h_text, h_gcz, h_gcx, w_text, w_gcz, w_gcx
        torch.cuda.empty_cache()
        return out

    # -- leg 2: the hg38-shaped genome, blocks LPT-sharded over the ranks -----------------------------------------------------------
    def leg_genome(self) -> dict:
        torch, G = self.torch, self.G
        from gecoz_b200 import sharding
        dev, lr, rank, world = self.dev, self.local_rank, self.rank, self.world
        self.wd.enter("cfg3: synthesis of this rank's blocks", 600)
        lengths, plan = genome_plan(self.args.scale)
        owner = sharding.lpt_assign([p[2] for p in plan], world)
        mine = [b for b, r in enumerate(owner) if r == rank]
        self.lengths, self.plan, self.owner, self.mine = lengths, plan, owner, mine
        t0 = time.perf_counter()
        h_text = {b: self.pinned(plan[b][2]) for b in mine}
        with ThreadPoolExecutor(max(1, min(len(mine), (os.cpu_count() or 8) // max(1, min(world, 8))))) as pool:
            list(pool.map(lambda b: block_text_into(h_text[b].numpy(), lengths, plan[b][1]), mine))
        synth_s = time.perf_counter() - t0
        d_text = {b: h_text[b].to(dev, non_blocking=True) for b in mine}
        shapes = {b: G.shape_from_counts(G.symbol_counts(d_text[b], lr)) for b in mine}
        gcx_len = {b: G.index_size(plan[b][2], 5) for b in mine}
        # every body of this rank stays on the device (they seed the replicated query index) and has a pinned host twin
        d_gcz = {b: torch.empty(int(shapes[b].size), dtype=torch.uint8, device=dev) for b in mine}
        d_gcx = {b: torch.empty(gcx_len[b], dtype=torch.uint8, device=dev) for b in mine}
        h_gcz = {b: self.pinned(int(shapes[b].size)) for b in mine}
        h_gcx = {b: self.pinned(gcx_len[b]) for b in mine}
        my_bases = sum(plan[b][2] - len(plan[b][1]) for b in mine)
        total_bases = int(sum(lengths))
        log(rank, f"cfg3: blocks {mine} ({my_bases} bp of {total_bases}), synthesis {synth_s:.1f} s")

        def dev_step():
            infos = []
            for b in mine:
                t = G.build_block(lr, d_text[b], plan[b][2], 32, shapes[b], d_gcz[b], d_gcx[b])
                t["_n"] = plan[b][2]
                infos.append(t)
            return infos

        def stage(b):
            return G.shape_from_counts(G.symbol_counts(h_text[b], lr))

        def e2e_step():
            # GecozFileWriter over this rank's blocks, two in flight: block k + 1 is counted / uploaded while block k is built,
            # and block k's bodies travel to the host while block k + 1 is built
            pending = []
            for b in mine:
                while len(pending) >= 2:
                    pending.pop(0).result()
                shp = stage(b)
                pending.append(self.builders.submit(G.build_block, lr, h_text[b], plan[b][2], 32, shp, h_gcz[b], h_gcx[b]))
            for f in pending:
                f.result()

        self.wd.enter("cfg3: device-resident steps", 600)
        steps = self.steps if world > 1 else max(1, min(self.steps, 3))
        warm = self.warm if world > 1 else 1
        clocks = ClockSampler(lr)
        if rank == 0 and world > 1:
            clocks.start()
        self.timed(dev_step, warm)
        w0 = time.time()
        ms_total, infos = self.timed(dev_step, steps)
        clk = clocks.stop(w0, time.time()) if (rank == 0 and world > 1) else None
        ms_step = self.reduce(ms_total / steps, "max")
        my_dev_ms = ms_total / steps

        self.wd.enter("cfg3: e2e steps", 600)
        self.timed(e2e_step, 1)
        t0 = time.perf_counter()
        ms_e2e_total, _ = self.timed(e2e_step, steps)
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms_e2e = self.reduce(max(ms_e2e_total, wall_ms) / steps, "max")

        # byte parity of every block of this rank against the oracle's digests, on the bodies the e2e arm wrote
        self.wd.enter("cfg3: parity digests", 600)
        parity = {"checked": False, "why": "no golden digests (tests/golden/full_size_digests.json) for this scale"}
        g3 = self.gold.get("cfg3")
        bad = []
        if g3 and self.args.scale == 1.0:
            for b in mine:
                gb = g3["blocks"][b]
                assert gb["headers"] == plan[b][0] and gb["n"] == plan[b][2], f"block {b}: the plan differs from the oracle's merge"
                if sha(h_gcz[b].numpy()) != gb["gcz_body"] or sha(h_gcx[b].numpy()) != gb["gcx_body"]:
                    bad.append(b)
                if not (torch.equal(h_gcz[b], d_gcz[b].cpu()) and torch.equal(h_gcx[b], d_gcx[b].cpu())):
                    bad.append(-b - 1)
            nbad = int(self.reduce(float(len(bad)), "sum"))
            parity = {"checked": True, "ok": nbad == 0, "blocks": len(plan),
                      "against": "oracle digests of every block body, tests/golden/full_size_digests.json (tools/make_digests.py)"}
            if bad:
                raise AssertionError(f"rank {rank}: blocks {bad} differ from the oracle's digests")
        flat = [t for step in infos for t in step]
        step_alg = float(sum(11.0 * p[2] for p in plan)) + float(self.reduce(float(sum(int(shapes[b].size) + gcx_len[b] for b in mine)), "sum"))
        biggest = max(p[2] for p in plan)
        out = {
            "value": total_bases / 1e6 / (ms_step / 1e3), "unit": "Mbp/s", "ms_per_step": ms_step, "steps": steps, "warmup": warm,
            "bases": total_bases, "blocks": len(plan), "blocks_per_rank": [owner.count(r) for r in range(world)],
            "lpt_efficiency_bound": sum(p[2] for p in plan) / (world * max(sum(plan[b][2] for b in range(len(plan)) if owner[b] == r) for r in range(world))),
            "clocks": clk,
            "e2e": {"value": total_bases / 1e6 / (ms_e2e / 1e3), "unit": "Mbp/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(self.reduce(float(sum(plan[b][2] for b in mine)), "sum")),
                    "d2h_bytes_per_step": int(self.reduce(float(sum(int(shapes[b].size) + gcx_len[b] for b in mine)), "sum")),
                    "pipelining": "per rank, GecozFileWriter's schedule with two blocks in flight per GPU: the upload + histogram of block k + 1 "
                                  "and the copies of block k - 1's bodies overlap the kernels of block k; pinned host text in, pinned host bodies out"},
            "gpu_launches": int(self.reduce(float(sum(t["kernel_launches"] for t in flat)), "sum")),
            "roofline": self.roofline_of(flat, biggest, my_dev_ms, step_alg, steps) if flat else None,
            "parity": parity, "synthesis_s": synth_s,
        }
        self.bodies = (d_gcz, d_gcx)
        self.block_text_dev = d_text
        del h_text, h_gcz, h_gcx
        return out

    # -- replicated index -------------------------------------------------------------------------------------------------------------
    def open_replicated(self):
        """Every rank ends up with a GSSA per block: the bodies of block b are broadcast by the rank that built it (NCCL), opened
        (re-laid out into rank sectors) and dropped."""
        torch, G, dist = self.torch, self.G, self.dist
        d_gcz, d_gcx = self.bodies
        self.wd.enter("replicate + open the index", 600)
        sizes = {b: (int(d_gcz[b].numel()), int(d_gcx[b].numel())) for b in self.mine}
        if self.world > 1:
            every = [None] * self.world
            dist.all_gather_object(every, sizes)
            sizes = {k: v for d in every for k, v in d.items()}
        t0 = time.perf_counter()
        gssas, index_bytes = [], 0
        for b, (headers, _, n) in enumerate(self.plan):
            if self.owner[b] == self.rank:
                z, x = d_gcz.pop(b), d_gcx.pop(b)
            else:
                z = torch.empty(sizes[b][0], dtype=torch.uint8, device=self.dev)
                x = torch.empty(sizes[b][1], dtype=torch.uint8, device=self.dev)
            if self.world > 1:
                dist.broadcast(z, self.owner[b])
                dist.broadcast(x, self.owner[b])
            torch.cuda.synchronize()
            gssas.append(G.GSSA.open(self.local_rank, z, n, x, headers))
            index_bytes += sizes[b][0] + sizes[b][1]
            del z, x
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        self.gssas, self.index_bytes = gssas, index_bytes
        return time.perf_counter() - t0

    def device_patterns(self, count: int, seed: int, chunk: int = 4_000_000):
        """cfg4/cfg5 patterns drawn on the GPU: length uniform in [15, 100]; half i.i.d. ACGT, half windows of this rank's
        largest block (windows touching N or a separator are re-drawn, then left random).  List of (bytes, offsets) chunks."""
        torch = self.torch
        dev = self.dev
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        text = self.block_text_dev[max(self.mine, key=lambda b: self.plan[b][2])] if self.mine else None
        chunks = []
        if text is not None:
            n = text.numel()
            bad = (text == 78) | (text == 0)
            badcum = torch.cumsum(bad, 0, dtype=torch.int32)
            del bad
        done = 0
        while done < count:
            c = min(chunk, count - done)
            lens = torch.randint(15, 101, (c,), generator=g, device=dev, dtype=torch.int64)
            off = torch.zeros(c + 1, dtype=torch.int64, device=dev)
            torch.cumsum(lens, 0, out=off[1:])
            total = int(off[-1].item())
            code = torch.randint(0, 4, (total,), generator=g, device=dev, dtype=torch.uint8)
            data = 65 + 2 * code + 2 * (code >= 2).to(torch.uint8) + 11 * (code == 3).to(torch.uint8)       # A C G T
            del code
            if text is not None:
                take = torch.rand(c, generator=g, device=dev) < 0.5
                starts = torch.randint(0, max(1, n - 102), (c,), generator=g, device=dev, dtype=torch.int64)
                for _ in range(8):
                    last = torch.clamp(starts + lens - 1, max=n - 1)
                    dirty = take & ((badcum[last] - badcum[starts] + ((text[starts] == 78) | (text[starts] == 0)).to(torch.int32)) > 0)
                    if not bool(dirty.any()):
                        break
                    starts = torch.where(dirty, torch.randint(0, max(1, n - 102), (c,), generator=g, device=dev, dtype=torch.int64), starts)
                last = torch.clamp(starts + lens - 1, max=n - 1)
                take &= (badcum[last] - badcum[starts] + ((text[starts] == 78) | (text[starts] == 0)).to(torch.int32)) == 0
                pid = torch.repeat_interleave(torch.arange(c, device=dev), lens, output_size=total)
                src = starts[pid] + (torch.arange(total, device=dev) - off[pid])
                data = torch.where(take[pid], text[torch.clamp(src, max=n - 1)], data)
                del pid, src, take, starts, last
            chunks.append((data.contiguous(), off))
            done += c
        return chunks

    # -- leg 3: cfg4, count ---------------------------------------------------------------------------------------------------------
    def leg_count(self) -> dict:
        torch, G, dist = self.torch, self.G, self.dist
        from gecoz_b200 import sharding
        dev, rank, world = self.dev, self.rank, self.world
        self.wd.enter("cfg4: pattern synthesis", 600)
        npat = int(self.args.count_patterns)
        lo, hi = sharding.shard_bounds(npat, world)[rank]
        width = sharding.shard_bounds(npat, world)[0][1]                       # the widest shard (rank 0's)
        chunks = self.device_patterns(hi - lo, seed=5 + rank)
        h_chunks = [(torch.empty(d.numel(), dtype=torch.uint8, pin_memory=True).copy_(d),
                     torch.empty(o.numel(), dtype=torch.int64, pin_memory=True).copy_(o)) for d, o in chunks]
        pat_bytes = sum(int(d.numel()) for d, _ in chunks)
        d_tot = torch.zeros(width, dtype=torch.int64, device=dev)
        h_tot = torch.empty(width, dtype=torch.int64).pin_memory()
        parts = [torch.empty(width, dtype=torch.int64, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None
        h_all = torch.empty((world, width), dtype=torch.int64).pin_memory() if rank == 0 else None

        def run(sources, out):
            at = 0
            for data, off in sources:
                c = int(off.numel()) - 1
                G.count_totals(self.gssas, data, off, out[at:at + c])
                at += c

        def cdev():
            run(chunks, d_tot)
            if world > 1:
                dist.gather(d_tot, parts, dst=0)

        def chost():
            # host shard in (pinned), per-pattern totals of every rank back on rank 0's host
            run(h_chunks, d_tot)
            if world > 1:
                dist.gather(d_tot, parts, dst=0)
                if rank == 0:
                    h_all.copy_(torch.stack(parts), non_blocking=True)
            else:
                h_all[0].copy_(d_tot, non_blocking=True)
            torch.cuda.synchronize()

        self.wd.enter("cfg4: count steps", 900)
        csteps = 2
        self.timed(cdev, 1)
        cms, _ = self.timed(cdev, csteps)
        cms = self.reduce(cms / csteps, "max")
        local_sum = int(d_tot[:hi - lo].sum().item())
        self.timed(chost, 1)
        t0 = time.perf_counter()
        cms_e2e, _ = self.timed(chost, csteps)
        cms_e2e = self.reduce(max(cms_e2e, (time.perf_counter() - t0) * 1e3) / csteps, "max")
        assert int(d_tot[:hi - lo].sum().item()) == local_sum, "host-buffer and device-buffer counts disagree"
        found = int(self.reduce(float((d_tot[:hi - lo] > 0).sum().item()), "sum"))
        # the gather moved every shard intact: per-rank checksums computed before the gather == checksums of the gathered rows
        if world > 1:
            sums = [None] * world
            dist.all_gather_object(sums, (hi - lo, local_sum))
            if rank == 0:
                for r, (cnt, s) in enumerate(sums):
                    assert int(h_all[r, :cnt].sum().item()) == s, f"gathered shard {r} differs from what rank {r} computed"
        # roofline of the backward search on the first chunk: rank sectors (32 B each) the kernels load / their device time
        # (events inside gcz_count_multi); the counters come from the same searches run once more with counting on
        c0 = int(chunks[0][1].numel()) - 1
        G.count_totals(self.gssas, chunks[0][0], chunks[0][1], d_tot[:c0])
        kms = G.last_query_stats()["kernel_ms"]
        st = G.count_stats(self.gssas, chunks[0][0], chunks[0][1])
        sector_gbs = st["rank_sectors"] * 32 / (kms / 1e3) / 1e9
        stats = {"bound": "hbm", "kernel": "count_kernel<1>: backward search of one chunk of the batch against every block (one launch per block)",
                 "achieved": sector_gbs, "peak": self.peak, "peak_source": self.peak_src, "unit": "GB/s", "frac": sector_gbs / self.peak,
                 "traffic": None, "patterns": c0, "launches": len(self.gssas), "kernel_ms": kms,
                 "algorithmic_bytes": st["rank_sectors"] * 32, "rank_sectors": st["rank_sectors"], "steps": st["steps"],
                 "reference_rank_calls": st["reference_rank_calls"],
                 "reference_layout": {"bytes": 74 * st["reference_rank_calls"], "GBps_equivalent": 74 * st["reference_rank_calls"] / (kms / 1e3) / 1e9,
                                      "what": "SURVEY.md 8(d) B_count: 74 B per RankedWTNode.count call of the reference's loop (64 B chunk + uint16 + uint64 "
                                              "counters); the re-laid-out index serves bit + rank from one 32 B sector and both ends of an interval from one load "
                                              "when they share it"},
                 "index_bytes": st["index_bytes"], "l2": "random 32 B sector reads; the index of one block (0.1-0.15 GB) is about the size of the 126 MB L2",
                 "interval_table": "the last K <= 12 symbols of a pattern are one lookup in a table of the block's K-symbol strings (built at open by "
                                   "the search itself; a lookup counts as one sector): rank_sectors / steps-with-table shrink ~3x against the "
                                   "reference's loop (reference_rank_calls, steps: counted without the table), so the kernel is latency- rather than "
                                   "bandwidth-bound and frac is lower than the 0.61 it had when every step loaded its sectors"}
        out = {"metric": "count queries/s against the hg38-shaped index (every pattern x every block, backward-search intervals summed)",
               "value": npat / (cms / 1e3), "unit": "queries/s", "patterns": npat, "blocks": len(self.gssas),
               "pattern_length": "uniform 15..100, 50% text-sampled (N-free windows of the rank's largest block) / 50% random",
               "found_somewhere": found, "ms_per_batch": cms, "index_bytes": self.index_bytes,
               "sharding": f"{world} contiguous shard(s), replicated index" + (", one NCCL gather of the per-pattern totals to rank 0" if world > 1 else ""),
               "e2e": {"value": npat / (cms_e2e / 1e3), "unit": "queries/s", "ms_per_batch": cms_e2e,
                       "h2d_bytes_per_step": int(self.reduce(float(pat_bytes + 8 * (hi - lo + len(chunks))), "sum")),
                       "d2h_bytes_per_step": int(npat * 8)}}
        if stats:
            out["roofline"] = stats
        self.count_chunks = chunks
        return out

    # -- leg 4: cfg5, locate --------------------------------------------------------------------------------------------------------
    def leg_locate(self) -> dict:
        torch, G = self.torch, self.G
        from gecoz_b200 import sharding
        rank, world = self.rank, self.world
        self.wd.enter("cfg5: locate", 900)
        npat = int(self.args.locate_patterns)
        lo, hi = sharding.shard_bounds(npat, world)[rank]
        mine = hi - lo
        # the shard = the first `mine` patterns of this rank's count batch (already on the device), as pinned host arrays
        srcs, left = [], mine
        for data, off in self.count_chunks:
            if left <= 0:
                break
            c = min(left, int(off.numel()) - 1)
            o = off[:c + 1].cpu()
            srcs.append((data[:int(o[-1])].cpu().pin_memory(), o.pin_memory()))
            left -= c
        self.barrier()
        t0 = time.perf_counter()
        occ = 0
        for data, off in srcs:
            occ += G.find_total(self.gssas, data, off)
        torch.cuda.synchronize()
        sec = self.reduce(time.perf_counter() - t0, "max")
        occ = int(self.reduce(float(occ), "sum"))
        return {"metric": "locate (GSSA.find: occurrences located through the sampled SA, sorted and split per string) against the hg38-shaped index",
                "value": npat / sec, "unit": "patterns/s", "patterns": npat, "occurrences": occ, "occurrences_per_s": occ / sec,
                "ms": sec * 1e3, "blocks": len(self.gssas), "timing": "wall clock of the calls, pinned host patterns in, host results out, max over ranks",
                "sharding": f"{world} contiguous shard(s), replicated index, results stay on the rank that computed them"}

    # -- CPU baseline (rank 0, N = 1) ---------------------------------------------------------------------------------------------------
    def cpu_baseline(self) -> dict:
        from gecoz_b200 import synth
        from oracle import gcz_oracle as O
        self.wd.enter("cpu baseline (oracle)", 900)
        O.build()
        text = synth.cfg2_text(self.args.length or CFG2_LEN, seed=3)
        t0 = time.perf_counter()
        O.build_block(text, 32, threads=2)
        sec = time.perf_counter() - t0
        return {"value": (len(text) - 1) / 1e6 / sec, "unit": "Mbp/s", "cores": 2, "kind": "port",
                "sample": f"the full {len(text) - 1} bp chr1-shaped block (the GPU arm's text), one run ({sec:.1f} s); C restatement of the "
                          f"Java path (no JVM on the box); one block can use 2 threads (SA-IS, then HSWT || SSA); host: {host_info()}"}

    # -- the run ----------------------------------------------------------------------------------------------------------------------
    def run(self) -> None:
        args, rank, world = self.args, self.rank, self.world
        line = {"metric": METRIC, "unit": "Mbp/s", "n_gpus": world, "steps": self.steps, "warmup": self.warm, "higher_is_better": True,
                "vs_baseline": None, "dtype": "u8/int32 (64-bit packed keys)", "data": "synthetic", "config": config_of(world, args.scale)}
        genome = None
        if world == 1:
            blk = self.leg_block()
            line.update({"scaling": "weak", **blk})
            if not args.block_only:
                genome = self.leg_genome()
                line["genome"] = genome
        else:
            genome = self.leg_genome()
            line.update({"scaling": "strong", **{k: genome[k] for k in ("value", "ms_per_step", "clocks", "e2e", "gpu_launches", "roofline", "parity")}})
            line["genome"] = {k: genome[k] for k in ("bases", "blocks", "blocks_per_rank", "lpt_efficiency_bound", "synthesis_s")}
        line["count"] = line["locate"] = None
        if genome is not None and not args.no_queries:
            open_s = self.open_replicated()
            line["count"] = self.leg_count()
            line["count"]["index_open_s"] = open_s
            line["locate"] = self.leg_locate()
            for g in self.gssas:
                g.close()
        line["cpu_baseline"] = self.cpu_baseline() if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
        self.wd.cancel()
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            self.wd.enter("shutdown", 120)
            self.dist.barrier()
            self.dist.destroy_process_group()
            self.wd.cancel()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--length", type=int, default=None, help="override the cfg2 block length (debugging)")
    ap.add_argument("--scale", type=float, default=1.0, help="hg38 sequence lengths x scale (debugging)")
    ap.add_argument("--count-patterns", type=int, default=100_000_000, help="cfg4 batch (whole job)")
    ap.add_argument("--locate-patterns", type=int, default=10_000_000, help="cfg5 batch (whole job)")
    ap.add_argument("--block-only", action="store_true", help="N = 1: only the cfg2 block legs")
    ap.add_argument("--no-queries", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-budget", type=float, default=1200.0, help="--impl reference: stop starting new steps after this many seconds")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    try:
        if args.impl == "reference":
            run_reference(args, rank, world)
        else:
            Bench(args).run()
    except BaseException as ex:                                   # loud: this rank's traceback, then a non-zero status for the whole job
        if isinstance(ex, SystemExit) and ex.code in (0, None):
            raise
        print(f"[bench rank {rank}] FAILED: {type(ex).__name__}: {ex}", file=sys.stderr, flush=True)
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)


if __name__ == "__main__":
    main()
