"""The native host layer (include/gcz_file.h) from Python: FASTA records, block planning, writer, reader.

The Python classes in gecoz_file.py / geco_index.py mirror the reference's Java classes one to one and are what the
parity tests were written against; this module drives the C++ implementation of the same host logic
(csrc/host_file.cpp), which is what a non-Python caller links.  Both produce the same bytes.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Sequence

import numpy as np

from . import _native as N
from .gssa import GSSA


def _cstr_array(strings: Sequence[str]):
    arr = (C.c_char_p * max(len(strings), 1))()
    for i, s in enumerate(strings):
        arr[i] = s.encode("latin-1")
    return arr


class Fasta:
    """FastaIterator with lazy = true (fasta/FastaIterator.java:39-127) over a plain file or a byte buffer."""

    def __init__(self, source):
        self._h = C.c_void_p()
        self._keep = None
        if isinstance(source, (str, Path)):
            N.check(N.lib().gcz_fasta_open(str(source).encode(), C.byref(self._h)))
        else:
            self._keep = np.frombuffer(bytes(source), dtype=np.uint8) if not isinstance(source, np.ndarray) else np.ascontiguousarray(source)
            N.check(N.lib().gcz_fasta_open_buffer(N.ptr(self._keep), len(self._keep), C.byref(self._h)))

    def __len__(self) -> int:
        return int(N.lib().gcz_fasta_count(self._h))

    def record(self, i: int) -> tuple[str, int, int, bool]:
        h, pos, ln, ml = C.c_char_p(), C.c_int64(), C.c_int64(), C.c_int32()
        N.check(N.lib().gcz_fasta_record(self._h, i, C.byref(h), C.byref(pos), C.byref(ln), C.byref(ml)))
        return h.value.decode("latin-1"), pos.value, ln.value, bool(ml.value)

    def read(self, i: int) -> np.ndarray:
        _, _, ln, _ = self.record(i)
        out = np.zeros(max(ln, 1), dtype=np.uint8)
        N.check(N.lib().gcz_fasta_read(self._h, i, N.ptr(out), len(out)))
        return out[:ln]

    def records(self):
        return [(self.record(i)[0], self.read(i)) for i in range(len(self))]

    def close(self):
        if self._h:
            N.lib().gcz_fasta_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def plan_blocks(lengths: Sequence[int], headers: Sequence[str]) -> list[list[int]]:
    """tools/GecoIndex.java:57-98: blocks in file order, each a list of sequence indices in block order."""
    n = len(lengths)
    ln = np.ascontiguousarray(lengths, dtype=np.int64)
    block_of, order = np.zeros(max(n, 1), np.int64), np.zeros(max(n, 1), np.int64)
    nb = int(N.lib().gcz_plan_blocks(N.ptr(ln), _cstr_array(headers), n, N.ptr(block_of), N.ptr(order)))
    N.check(nb)
    blocks: list[list[int]] = [[] for _ in range(nb)]
    for i in sorted(range(n), key=lambda i: order[i]):
        if block_of[i] >= 0:
            blocks[int(block_of[i])].append(i)
    return blocks


def ref_header(headers: Sequence[str], block_size: int, text_len: int) -> bytes:
    out = np.zeros(int(N.lib().gcz_ref_header_length(_cstr_array(headers), len(headers))), np.uint8)
    w = int(N.lib().gcz_ref_header_write(_cstr_array(headers), len(headers), block_size, text_len, N.ptr(out), len(out)))
    N.check(w)
    return out[:w].tobytes()


def ssa_header(headers: Sequence[str], index_len: int) -> bytes:
    out = np.zeros(25, np.uint8)
    N.lib().gcz_ssa_header_write(_cstr_array(headers), len(headers), index_len, N.ptr(out))
    return out.tobytes()


def header_hash(headers: Sequence[str]) -> int:
    return int(N.lib().gcz_header_hash(_cstr_array(headers), len(headers)))


def index(fasta: Fasta, gcz_path, gcx_path=None, sampling: int = 32, devices: Sequence[int] = (0,), engine: N.Engine | None = None) -> dict:
    """GecoIndex.index + GecozFileWriter in native code (gcz_index_fasta)."""
    devs = (C.c_int * len(devices))(*devices)
    rep = N.IndexReport()
    N.check(N.lib().gcz_index_fasta(fasta._h, str(gcz_path).encode(), None if gcx_path is None else str(gcx_path).encode(),
                                    sampling, len(devices), devs, C.byref(engine) if engine is not None else None, C.byref(rep)))
    return {"blocks": rep.blocks, "sequences": rep.sequences, "symbols": rep.symbols, "seconds": rep.seconds}


class Reader:
    """GecozFileReader (fmt/GecozFileReader.java:57-200) in native code."""

    def __init__(self, path):
        self._h = C.c_void_p()
        N.check(N.lib().gcz_reader_open(str(path).encode(), C.byref(self._h)))

    @property
    def n_blocks(self) -> int:
        return int(N.lib().gcz_reader_num_blocks(self._h))

    @property
    def sampling_factor(self) -> int:
        return int(N.lib().gcz_reader_sampling_factor(self._h))

    def block(self, b: int) -> dict:
        ln, size, nh = C.c_int64(), C.c_int64(), C.c_int32()
        N.check(N.lib().gcz_reader_block(self._h, b, C.byref(ln), C.byref(size), C.byref(nh)))
        return {"len": ln.value, "size": size.value,
                "headers": [N.lib().gcz_reader_header(self._h, b, i).decode("latin-1") for i in range(nh.value)]}

    def find(self, header: str) -> tuple[int, int]:
        b, s = C.c_int32(), C.c_int32()
        N.check(N.lib().gcz_reader_find(self._h, header.encode("latin-1"), C.byref(b), C.byref(s)))
        return b.value, s.value

    def open_block(self, b: int, device: int = 0) -> GSSA:
        h = C.c_void_p()
        N.check(N.lib().gcz_reader_open_block(self._h, b, device, C.byref(h)))
        return GSSA(h, device, self.block(b)["headers"])

    def _text(self, rc: int, out: C.c_void_p, n: C.c_int64) -> list[str]:
        N.check(rc)
        try:
            return C.string_at(out, n.value).decode("latin-1").splitlines()
        finally:
            N.lib().gcz_free(out)

    def match(self, header: str | None, pattern: bytes, with_positions: bool = True, device: int = 0,
              engine: N.QueryEngine | None = None) -> list[str]:
        """GecoMatch.match / count (tools/GecoMatch.java:51-157): the lines the tool prints."""
        out, n = C.c_void_p(), C.c_int64()
        pat = np.frombuffer(bytes(pattern), dtype=np.uint8)
        rc = N.lib().gcz_match(self._h, device, None if header is None else header.encode("latin-1"), N.ptr(pat), len(pat),
                               1 if with_positions else 0, C.byref(engine) if engine is not None else None, C.byref(out), C.byref(n))
        return self._text(rc, out, n)

    def gff_search(self, patterns: bytes, device: int = 0, engine: N.QueryEngine | None = None) -> list[str]:
        """SimpleGFFGenerator.search (tools/SimpleGFFGenerator.java:45-163) for the bytes of a pattern file."""
        out, n = C.c_void_p(), C.c_int64()
        data = np.frombuffer(bytes(patterns), dtype=np.uint8)
        rc = N.lib().gcz_gff_search(self._h, device, N.ptr(data) if len(data) else None, len(data),
                                    C.byref(engine) if engine is not None else None, C.byref(out), C.byref(n))
        return self._text(rc, out, n)

    def extract_fasta(self, path, device: int = 0, engine: N.QueryEngine | None = None) -> int:
        """GecoRead.fasta (tools/GecoRead.java:83-175)."""
        n = C.c_int64()
        N.check(N.lib().gcz_extract_fasta(self._h, device, str(path).encode(), C.byref(engine) if engine is not None else None, C.byref(n)))
        return n.value

    def extract_sequence(self, header: str, start: int, end: int, path, device: int = 0, engine: N.QueryEngine | None = None) -> int:
        """GecoRead.sequence (tools/GecoRead.java:33-81): symbols [start, min(end, length)) of one sequence into `path`."""
        n = C.c_int64()
        N.check(N.lib().gcz_extract_sequence(self._h, device, header.encode("latin-1"), start, end, str(path).encode(),
                                             C.byref(engine) if engine is not None else None, C.byref(n)))
        return n.value

    def close(self):
        if self._h:
            N.lib().gcz_reader_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
