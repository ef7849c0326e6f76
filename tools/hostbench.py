#!/usr/bin/env python
"""hostbench.py — the HOST side of `gecotools -o` (gcz_index_fasta) timed without a GPU.

    python tools/hostbench.py [--fasta PATH | --scale 1.0] [--out DIR] [--gpu-gbps 12.6]

The per-block device work is replaced by a stand-in engine (compiled here with gcc): count_symbols samples the text,
build_block sleeps for n / --gpu-gbps (the measured device throughput of a block build) and fills both output slices
(with n % 251 and n % 241).
What remains is what the host has to do around the GPU: scan the FASTA, assemble the block texts, create the files and
take the bodies.  The output files are NOT valid indexes; the tool reports seconds per stage.
"""
from __future__ import annotations

import argparse
import ctypes as C
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

ENGINE_C = r"""
#include <stdint.h>
#include <string.h>
#include <time.h>
typedef struct gcz_shape gcz_shape;
static double gbps = 12.6;
void standin_set_gbps(double v) { gbps = v; }
int standin_count(int device, const uint8_t* text, int64_t n, int64_t counts[256]) {
    (void)device;
    memset(counts, 0, 256 * sizeof(int64_t));
    for (int64_t i = 0; i < n; i += 64) counts[text[i]] += 64;       /* a plausible histogram, not the exact one */
    if (counts[0] == 0) counts[0] = 1;
    return 0;
}
int standin_build(int device, const uint8_t* text, int64_t n, int32_t rate, const gcz_shape* shape, uint8_t* gcz_body,
                  int64_t gcz_len, uint8_t* gcx_body, int64_t gcx_len, int32_t* sa, uint8_t* bwt) {
    (void)device; (void)text; (void)rate; (void)shape; (void)sa; (void)bwt;
    struct timespec ts;
    double s = (double)n / (gbps * 1e9);
    ts.tv_sec = (time_t)s;
    ts.tv_nsec = (long)((s - (double)ts.tv_sec) * 1e9);
    nanosleep(&ts, 0);
    memset(gcz_body, (int)(n % 251), (size_t)gcz_len);               /* recognisable per block: the tests check the files */
    memset(gcx_body, (int)(n % 241), (size_t)gcx_len);
    return 0;
}
"""


def standin_engine(gbps: float):
    from gecoz_b200 import _native as N
    d = Path(tempfile.mkdtemp(prefix="gcz_hostbench_"))
    (d / "engine.c").write_text(ENGINE_C)
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", str(d / "engine.so"), str(d / "engine.c")], check=True)
    so = C.CDLL(str(d / "engine.so"))
    so.standin_set_gbps.argtypes = [C.c_double]
    so.standin_set_gbps(gbps)
    eng = N.Engine(C.cast(so.standin_count, N.COUNT_SYMBOLS_FN), C.cast(so.standin_build, N.BUILD_BLOCK_FN))
    eng._keep = so
    return eng


def write_fasta(path: Path, scale: float, width: int = 60) -> None:
    from gecoz_b200 import synth
    with open(path, "wb") as f:
        for i, (name, ln) in enumerate(zip(synth.HG38_NAMES, synth.HG38_LENGTHS)):
            ln = max(8, int(ln * scale))
            seq = synth.chromosome_shaped(ln, 4 + i)
            f.write(b">" + name.encode() + b"\n")
            full = (ln // width) * width
            a = np.empty((full // width, width + 1), np.uint8)
            a[:, :width] = seq[:full].reshape(-1, width)
            a[:, width] = 10
            f.write(a.tobytes())
            if ln > full:
                f.write(seq[full:].tobytes() + b"\n")


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--fasta", default=None)
    ap.add_argument("--scale", type=float, default=1.0, help="hg38-shaped synthetic FASTA of this scale when --fasta is not given")
    ap.add_argument("--out", default="/dev/shm")
    ap.add_argument("--gpu-gbps", type=float, default=12.6, help="device throughput the stand-in build sleeps for (Gbp/s)")
    ap.add_argument("--devices", type=int, default=1)
    ap.add_argument("--engine", default="standin", choices=["standin", "cuda"], help="cuda: the library's own entry points (needs a GPU); "
                    "the files are then real indexes and their sha256 is printed")
    ap.add_argument("--repeat", type=int, default=1, help="index the same FASTA this many times in one process: the first run pays for the CUDA "
                    "context, the module load and the device arena; the later ones are what a long-lived process sees")
    args = ap.parse_args()
    from gecoz_b200 import native_file as NF
    out = Path(args.out)
    fasta = Path(args.fasta) if args.fasta else out / f"hostbench_{args.scale:g}.fa"
    if not fasta.exists():
        t = time.perf_counter()
        write_fasta(fasta, args.scale)
        print(f"wrote {fasta} ({fasta.stat().st_size / 1e9:.2f} GB) in {time.perf_counter() - t:.1f} s")
    eng = standin_engine(args.gpu_gbps) if args.engine == "standin" else None
    t = time.perf_counter()
    f = NF.Fasta(fasta)
    t_scan = time.perf_counter() - t
    total = sum(f.record(i)[2] for i in range(len(f)))
    for r in range(max(1, args.repeat)):
        t = time.perf_counter()
        rep = NF.index(f, out / "hostbench.gcz", None, 32, tuple(range(args.devices)), eng)
        t_index = time.perf_counter() - t
        if args.repeat > 1:
            print(f"run {r + 1}: index {t_index:7.3f} s = {total / 1e6 / t_index:.0f} Mbp/s")
    f.close()
    dev = total / (args.gpu_gbps * 1e9) / args.devices
    print(f"records {rep['sequences']}  blocks {rep['blocks']}  symbols {rep['symbols']}")
    print(f"scan   {t_scan:7.3f} s  ({fasta.stat().st_size / 1e9 / t_scan:.2f} GB/s)")
    what = "stand-in device time" if args.engine == "standin" else f"device time at {args.gpu_gbps} Gbp/s would be"
    print(f"index  {t_index:7.3f} s  ({what} {dev:.3f} s on {args.devices} device(s): host overhead {t_index - dev:.3f} s)")
    print(f"total  {t_scan + t_index:7.3f} s  = {total / 1e6 / (t_scan + t_index):.0f} Mbp/s wall")
    for p in (out / "hostbench.gcz", out / "hostbench.gcx"):
        if args.engine == "cuda":
            import hashlib
            h = hashlib.sha256()
            with open(p, "rb") as fh:
                while chunk := fh.read(1 << 24):
                    h.update(chunk)
            print(f"{p.name}: {p.stat().st_size} bytes, sha256 {h.hexdigest()}")
        p.unlink(missing_ok=True)


if __name__ == "__main__":
    main()
