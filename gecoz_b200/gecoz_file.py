"""`.gcz` / `.gcx` containers: the host-side mirror of nova-formats' gecoz package.

Same names and semantics as the Java classes (there is no JVM in this environment, so the host layer above
the C ABI is Python; the Java shim a maintainer would use instead is shown in INTEGRATION.md):

  GecozRefBlockHeader   fmt/GecozRefBlockHeader.java:36-137
  GecozSSABlockHeader   fmt/GecozSSABlockHeader.java:36-79
  GecozFileWriter       fmt/GecozFileWriter.java:60-310   (BlockWriter.run -> one gcz_build_block call)
  GecozFileReader       fmt/GecozFileReader.java:57-200   (read -> gcz_open_block -> GSSA)

(fmt/ = /root/reference/java/nova-formats/src/main/java/es/elixir/bsc/ngs/nova/gecoz/)
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from concurrent.futures import Future, ThreadPoolExecutor
from pathlib import Path
from typing import Iterable, Sequence

import numpy as np

from . import _native as N
from .gssa import GSSA

_MASK64 = (1 << 64) - 1


def _to_signed64(v: int) -> int:
    v &= _MASK64
    return v - (1 << 64) if v >= (1 << 63) else v


class GecozRefBlockHeader:
    MAGIC = b"GecozBWT"
    version = 1

    def __init__(self, headers: Sequence[str], size: int, length: int):
        self.headers = list(headers)
        self.size = int(size)       # the block size (header + shape table + nodes)
        self.len = int(length)      # the length of the generalized string

    # GecozRefBlockHeader(InputStream)  :59-82 — magic/version are read but never validated there
    @classmethod
    def parse(cls, buf: bytes, pos: int = 0) -> "GecozRefBlockHeader":
        size = int.from_bytes(buf[pos + 9:pos + 17], "little", signed=True)
        length = int.from_bytes(buf[pos + 17:pos + 25], "little", signed=True)
        headers, p = [], pos + 25
        while p < len(buf) and buf[p] > 0:
            q = buf.index(b"\0", p)
            headers.append(buf[p:q].decode("latin-1"))
            p = q + 1
        return cls(headers, size, length)

    def to_bytes(self) -> bytes:                                  # write(ByteBuffer) :90-101
        out = bytearray(self.MAGIC)
        out.append(self.version)
        out += self.size.to_bytes(8, "little", signed=True)
        out += self.len.to_bytes(8, "little", signed=True)
        for h in self.headers:
            out += h.encode("latin-1") + b"\0"
        out.append(0)
        return bytes(out)

    def findHeader(self, header: str) -> int:                     # :103-110
        for i, h in enumerate(self.headers):
            if header == h:
                return i
        return -1

    def getBlockHeaderLength(self) -> int:
        return self.block_header_length(self.headers)

    def getHeaderHash(self) -> int:
        return self.block_header_hash(self.headers)

    @staticmethod
    def block_header_hash(headers: Iterable[str]) -> int:         # getBlockHeaderHash :120-128
        h = 1125899906842597
        for header in headers:
            for ch in header:
                h = ((h << 5) - h + ord(ch)) & _MASK64
        return _to_signed64(h)

    @staticmethod
    def block_header_length(headers: Iterable[str]) -> int:       # getBlockHeaderLength :130-136
        return 26 + sum(len(h) + 1 for h in headers)


class GecozSSABlockHeader:
    MAGIC = b"GecozSSA"
    version = 1
    LENGTH = 25

    def __init__(self, headers: Sequence[str] | None, length: int, hash_: int | None = None):
        self.len = int(length)
        self.hash = GecozRefBlockHeader.block_header_hash(headers) if hash_ is None else int(hash_)

    @classmethod
    def parse(cls, buf: bytes, pos: int = 0) -> "GecozSSABlockHeader":     # :52-67
        if len(buf) - pos < 25:
            raise EOFError("truncated GecozSSA header")
        if buf[pos:pos + 8] != cls.MAGIC or buf[pos + 8] != cls.version:
            raise N.GczFormatError(N.GCZ_E_FORMAT, "bad GecozSSA magic")
        return cls(None, int.from_bytes(buf[pos + 9:pos + 17], "little", signed=True),
                   int.from_bytes(buf[pos + 17:pos + 25], "little", signed=True))

    def to_bytes(self) -> bytes:                                  # write :69-74
        return (self.MAGIC + bytes([self.version]) + self.len.to_bytes(8, "little", signed=True)
                + self.hash.to_bytes(8, "little", signed=True))


def ssa_path_for(ref_path: Path) -> Path:
    """`x.gcz` -> `x.gcx`, anything else -> name + 'gcx'  (fmt/GecozFileWriter.java:97-104)."""
    name = ref_path.name
    if name.endswith(".gcz"):
        name = name[:-3]
    return ref_path.with_name(name + "gcx")


def symbol_counts(text, device: int = 0) -> np.ndarray:
    """The counting loop of GecozFileWriter.write (:127-130), on the GPU."""
    counts = np.zeros(256, dtype=np.int64)
    n = text.numel() if hasattr(text, "numel") else len(text)
    N.check(N.lib().gcz_count_symbols(device, N.ptr(text), n, N.ptr(counts)))
    return counts


def shape_from_counts(counts: np.ndarray) -> N.Shape:
    s = N.Shape()
    c = np.ascontiguousarray(counts, dtype=np.int64)
    N.check(N.lib().gcz_shape_from_counts(N.ptr(c), C.byref(s)))
    return s


def index_size(n: int, sampling_factor: int) -> int:
    return int(N.lib().gcz_index_size(n, sampling_factor))


def build_block(device: int, text, n: int, sampling_rate: int, shape: N.Shape, gcz_body, gcx_body,
                sa_out=None, bwt_out=None) -> dict:
    """BlockWriter.run (fmt/GecozFileWriter.java:256-284) as one native call; returns the device timing."""
    sf = sampling_rate.bit_length() - 1
    N.check(N.lib().gcz_build_block(device, N.ptr(text), n, sampling_rate, C.byref(shape),
                                    N.ptr(gcz_body), int(shape.size), N.ptr(gcx_body), index_size(n, sf),
                                    N.ptr(sa_out), N.ptr(bwt_out)))
    t = N.BuildTiming()
    N.lib().gcz_last_build_timing(C.byref(t))
    return t.as_dict()


class _BodyWriter:
    """Where block bodies go before they reach the files: reusable host buffers, written with pwrite on helper threads.

    The reference hands BlockWriter two mapped file slices.  A mapping takes a page fault per 4 KiB while the build call
    fills it (125 ms per 128 MB measured) and the next block waits behind that; a buffer that is reused does
    not fault, and its pwrite overlaps the next build.  One buffer holds a whole block: ref header | .gcz body, then
    ssa header | .gcx body, so each file gets one write per block."""

    def __init__(self, max_buffers: int, threads: int = 2):
        self._pool = ThreadPoolExecutor(max_workers=threads)
        self._cv = threading.Condition()
        self._free: list[np.ndarray] = []
        self._out, self._max = 0, max(1, max_buffers)
        self._pending: list[Future] = []

    def take(self, nbytes: int) -> np.ndarray:
        with self._cv:
            self._cv.wait_for(lambda: self._out < self._max)
            self._out += 1
            best = None
            for k, b in enumerate(self._free):
                if len(b) >= nbytes and (best is None or len(b) < len(self._free[best])):
                    best = k
            if best is not None:
                return self._free.pop(best)
            self._free.clear()                                    # all too small
        return np.empty(max(nbytes, 1), dtype=np.uint8)

    def give(self, buf: np.ndarray) -> None:
        with self._cv:
            self._free.append(buf)
            self._out -= 1
            self._cv.notify_all()

    def commit(self, buf: np.ndarray, writes: Sequence[tuple[int, int, int, int]]) -> None:
        """writes: (fd, file offset, buffer offset, length); the buffer comes back to the pool when they are done."""
        fut = self._pool.submit(self._write, buf, list(writes))
        with self._cv:
            self._pending.append(fut)

    def _write(self, buf, writes):
        try:
            mv = memoryview(buf)
            for fd, off, at, length in writes:
                done = 0
                while done < length:
                    done += os.pwrite(fd, mv[at + done:at + length], off + done)
        finally:
            self.give(buf)

    def drain(self) -> None:
        with self._cv:
            pending, self._pending = self._pending, []
        for f in pending:
            f.result()

    def close(self) -> None:
        try:
            self.drain()
        finally:
            self._pool.shutdown(wait=True)


def _block_layout(header_len: int, ref_len: int, ssa_len: int) -> tuple[int, int, int]:
    """Offsets of the .gcz body and the .gcx body in a block buffer (64-byte aligned, their headers just before), and its size."""
    ref_at = (header_len + 63) & ~63
    ssa_at = ((ref_at + ref_len + 63) & ~63) + 64
    return ref_at, ssa_at, ssa_at + ssa_len


class _CudaEngine:
    """The CUDA path behind the C ABI (the only engine the product uses; the CPU tests inject the oracle here)."""

    def symbol_counts(self, text, device):
        return symbol_counts(text, device)

    def build_block(self, device, text, n, sampling_rate, shape, gcz_out, gcx_out):
        return build_block(device, text, n, sampling_rate, shape, gcz_out, gcx_out)


class GecozFileWriter:
    """Writes blocks to `ref_path` (.gcz) and `ssa_path` (.gcx).

    `devices` replaces the reference's `threads`: two blocks in flight per GPU — one being built, the next one
    being counted, which is also its upload (gcz_count_symbols leaves the text on the device for the build) — where
    the reference runs up to `-t` BlockWriters on a JDK pool.  File offsets are fixed in write() before the block
    is computed, exactly as in the reference, so blocks may finish in any order.
    """

    def __init__(self, ref_path, ssa_path=None, sampling_rate: int = 32, devices: Sequence[int] = (0,), engine=None):
        self.ref_path = Path(ref_path)
        self.ssa_path = Path(ssa_path) if ssa_path is not None else ssa_path_for(self.ref_path)
        if sampling_rate <= 0 or sampling_rate & (sampling_rate - 1):
            raise ValueError("sampling rate must be a power of two")
        self.sampling_rate = sampling_rate
        self.devices = list(devices)
        self._engine = engine or _CudaEngine()
        self._ref = open(self.ref_path, "w+b")
        self._ssa = open(self.ssa_path, "w+b")
        self._ref_pos = 0
        self._ssa_pos = 0
        self._pool = ThreadPoolExecutor(max_workers=2 * len(self.devices))
        self._free = list(self.devices) * 2               # two tokens per device: the library has two text slots
        self._cv = threading.Condition()
        self._jobs: list[Future] = []
        self._bodies = _BodyWriter(max_buffers=2 * len(self.devices) + 2)
        self.timings: list[dict] = []

    def write(self, headers: Sequence[str], text) -> None:
        """GecozFileWriter.write(String[] headers, ByteBuffer in)  :124-159."""
        text = np.ascontiguousarray(np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else text)
        n = len(text)
        device = self._acquire()
        try:
            counts = self._engine.symbol_counts(text, device)
            shape = shape_from_counts(counts)
            hdr = GecozRefBlockHeader(headers, GecozRefBlockHeader.block_header_length(headers) + shape.size, n)
            idx_size = index_size(n, self.sampling_rate.bit_length() - 1)
            ref_pos, ssa_pos = self._ref_pos, self._ssa_pos
            self._ref_pos += hdr.size
            self._ssa_pos += GecozSSABlockHeader.LENGTH + idx_size
            hb = hdr.to_bytes()
            sb = GecozSSABlockHeader(headers, idx_size).to_bytes()
        except BaseException:
            self._release(device)
            raise
        fut = self._pool.submit(self._block_writer, device, text, n, shape, hb, ref_pos, int(shape.size), sb, ssa_pos, idx_size)
        self._jobs.append(fut)

    # BlockWriter.run :256-284 — plus the retry contract of WriterPoolExecutor.afterExecute :203-226
    def _block_writer(self, device, text, n, shape, hb, ref_pos, ref_len, sb, ssa_pos, ssa_len):
        try:
            ref_at, ssa_at, size = _block_layout(len(hb), ref_len, ssa_len)
            buf = self._bodies.take(size)
            try:
                attempts = 0
                while True:
                    try:
                        t = self._engine.build_block(device, text, n, self.sampling_rate, shape, buf[ref_at:ref_at + ref_len],
                                                     buf[ssa_at:ssa_at + ssa_len])
                        break
                    except N.GczOutOfMemory:
                        attempts += 1
                        if attempts > 1:
                            raise
                        with self._cv:                   # wait until nothing else is in flight, then retry once
                            self._cv.wait_for(lambda: len(self._free) == 2 * len(self.devices) - 1, timeout=600)
            except BaseException:
                self._bodies.give(buf)
                raise
            t = dict(t or {})
            t["n"] = n
            self.timings.append(t)
        finally:
            self._release(device)                        # the files are written after the device is free again
        buf[ref_at - len(hb):ref_at] = np.frombuffer(hb, np.uint8)
        buf[ssa_at - len(sb):ssa_at] = np.frombuffer(sb, np.uint8)
        self._bodies.commit(buf, [(self._ref.fileno(), ref_pos, ref_at - len(hb), len(hb) + ref_len),
                                  (self._ssa.fileno(), ssa_pos, ssa_at - len(sb), len(sb) + ssa_len)])

    def _acquire(self) -> int:
        with self._cv:
            self._cv.wait_for(lambda: bool(self._free))
            return self._free.pop(0)

    def _release(self, device: int) -> None:
        with self._cv:
            self._free.append(device)
            self._cv.notify_all()

    def close(self) -> None:                              # :161-172
        try:
            for f in self._jobs:
                f.result()
        finally:
            try:
                self._pool.shutdown(wait=True)
                self._bodies.close()                      # the pending writes finish while the files are open
            finally:
                self._ref.close()
                self._ssa.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class GecozFileReader:
    """fmt/GecozFileReader.java:57-200.  Blocks are opened on `device`."""

    def __init__(self, path, device: int = 0):
        self.path = Path(path)
        self.device = device
        self._ref = np.memmap(self.path, dtype=np.uint8, mode="r")
        ssa_path = ssa_path_for(self.path)
        self._ssa = np.memmap(ssa_path, dtype=np.uint8, mode="r") if ssa_path.is_file() and os.access(ssa_path, os.R_OK) else None
        self.headers: dict[GecozRefBlockHeader, int] = {}
        position, total = 0, len(self._ref)
        while True:                                                         # :79-89 (do/while)
            head = bytes(self._ref[position:position + min(total - position, 1 << 20)])
            header = GecozRefBlockHeader.parse(head)
            self.headers[header] = position
            position += header.size
            if header.size <= 0 or position >= total:
                break

    def findBlockHeader(self, header: str):
        for h in self.headers:
            if h.findHeader(header) >= 0:
                return h
        return None

    def getBlockHeaders(self):
        return list(self.headers.keys())

    def read(self, header: GecozRefBlockHeader) -> GSSA | None:               # :115-177
        pos = self.headers.get(header)
        if pos is None:
            return None
        if self._ssa is None:
            # the reference returns a GSSA that can never locate; fail fast instead
            raise N.GczError(N.GCZ_E_ARG, f"{ssa_path_for(self.path)} is missing: queries need the .gcx index")
        hlen = header.getBlockHeaderLength()
        body = self._ref[pos + hlen:pos + header.size]
        # recover the sampling factor from the .gcx size  :134-149
        ssa_data_length = len(self._ssa) - len(self.headers) * GecozSSABlockHeader.LENGTH
        sf = -1
        while True:
            sf += 1
            if sf > 30:
                raise N.GczFormatError(N.GCZ_E_FORMAT, "invalid index file")
            if ssa_data_length >= sum(index_size(h.len, sf) for h in self.headers):
                break
        ssa_pos = 0
        for h in self.headers:
            if h is header:
                break
            ssa_pos += GecozSSABlockHeader.LENGTH + index_size(h.len, sf)
        ssa_size = index_size(header.len, sf)
        ssa_header = GecozSSABlockHeader.parse(bytes(self._ssa[ssa_pos:ssa_pos + 25]))
        if header.getHeaderHash() != ssa_header.hash or ssa_header.len != ssa_size:
            raise N.GczFormatError(N.GCZ_E_FORMAT, "invalid index file")     # :165-172
        idx = self._ssa[ssa_pos + 25:ssa_pos + 25 + ssa_size]
        return GSSA.open(self.device, body, header.len, idx, headers=header.headers)

    def close(self):
        self._ref = None
        self._ssa = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @staticmethod
    def checkFormat(path) -> bool:                                              # :190-199
        with open(path, "rb") as f:
            return f.read(8) == GecozRefBlockHeader.MAGIC
