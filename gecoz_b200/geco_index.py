"""Index driver: FASTA -> chromosome-bounded blocks -> GecozFileWriter, blocks scheduled over GPUs.

Mirror of tools/GecoIndex.java:51-146, fmt/GecozRefBlock.java:38-71 and fasta/TFastaSequence.java:46-52
(/root/reference/java/nova-gecoz/.../tools/, nova-formats/.../gecoz/, nova-formats/.../fasta/).  Which
sequences share a block, their order inside it and the order of blocks in the file all define file bytes,
so the merge below follows the reference's TreeSet semantics step by step.
"""
from __future__ import annotations

import bisect
import functools
import gzip
import time
from pathlib import Path
from typing import Iterable, Iterator, Sequence

import numpy as np

from .gecoz_file import GecozFileWriter


@functools.total_ordering
class FastaSequence:
    """fasta/TFastaSequence.java: ordered by length DESCENDING, then header ascending (:46-52)."""

    __slots__ = ("header", "length", "sequence", "id")

    def __init__(self, header: str, length: int, sequence=None, id_: int = 0):
        self.header, self.length, self.sequence, self.id = header, int(length), sequence, id_

    def _key(self):
        return (-self.length, self.header)

    def __eq__(self, o):
        return self._key() == o._key()

    def __lt__(self, o):
        return self._key() < o._key()


class GecozRefBlock:
    """fmt/GecozRefBlock.java: a TreeSet of sequences plus a running size that counts duplicates too."""

    def __init__(self, sequence: FastaSequence):
        self.sequences: list[FastaSequence] = [sequence]
        self.size = _int32(sequence.length + 1)

    def add(self, sequence: FastaSequence) -> None:                          # :45-48
        i = bisect.bisect_left(self.sequences, sequence)
        if i == len(self.sequences) or self.sequences[i] != sequence:       # a TreeSet drops equal elements
            self.sequences.insert(i, sequence)
        self.size = _int32(self.size + sequence.length + 1)

    def add_all(self, sequences: Iterable[FastaSequence]) -> None:           # :50-54
        for s in list(sequences):
            self.add(s)

    def compare(self, o: "GecozRefBlock") -> int:                            # compareTo :62-69
        if self.size != o.size:
            return 1 if self.size > o.size else -1
        a, b = self.sequences[0], o.sequences[0]
        return 0 if a == b else (-1 if a < b else 1)


def _int32(v: int) -> int:
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v >= (1 << 31) else v


def _tree_add(blocks: list, block, cmp) -> bool:
    """TreeSet.add with an explicit comparator: ordered insert, equal elements are rejected."""
    lo, hi = 0, len(blocks)
    while lo < hi:
        mid = (lo + hi) // 2
        c = cmp(block, blocks[mid])
        if c == 0:
            return False
        if c < 0:
            hi = mid
        else:
            lo = mid + 1
    blocks.insert(lo, block)
    return True


def merge_blocks(sequences: Sequence[FastaSequence]) -> list[GecozRefBlock]:
    """tools/GecoIndex.java:57-98: one block per sequence, greedy merge of the two smallest blocks while
    the sum does not exceed the largest original block, then order by longest member (desc)."""
    by_size = lambda a, b: a.compare(b)
    blocks: list[GecozRefBlock] = []
    for s in sequences:
        _tree_add(blocks, GecozRefBlock(s), by_size)
    if not blocks:
        return []
    max_size = blocks[-1].size                                               # :72
    while len(blocks) > 1:                                                   # :73-85
        first, second = blocks.pop(0), blocks.pop(0)
        size = _int32(first.size + second.size)
        if 0 < size <= max_size:
            first.add_all(second.sequences)
            _tree_add(blocks, first, by_size)
        else:
            _tree_add(blocks, first, by_size)
            _tree_add(blocks, second, by_size)
            break

    def by_longest(a: GecozRefBlock, b: GecozRefBlock) -> int:               # :88-96
        la, lb = a.sequences[0].length, b.sequences[0].length
        if la != lb:
            return -1 if la > lb else 1
        return a.compare(b)

    ordered: list[GecozRefBlock] = []
    for b in blocks:
        _tree_add(ordered, b, by_longest)
    return ordered


# ---- FASTA (host I/O; stays on the CPU like nova-gzip / fasta in the reference) -------------------------------
def read_fasta(path) -> Iterator[tuple[str, bytes]]:
    """Records as the reference sees them (fasta/FastaIterator.java:39-127, fasta/FastaFileReader.java:71-160): the
    native record scanner (csrc/host_file.cpp) is the one implementation of that state machine; gzipped files are
    decompressed on the host (nova-gzip's job in the reference), by the scanner's zlib or, failing that, here."""
    from . import _native as N
    from .native_file import Fasta
    path = Path(path)
    try:
        fasta = Fasta(path)
    except N.GczFormatError:
        fasta = Fasta(gzip.open(path, "rb").read())
    with fasta:
        for i in range(len(fasta)):
            yield fasta.record(i)[0], fasta.read(i).tobytes()


def block_text(block: GecozRefBlock) -> tuple[list[str], np.ndarray]:
    """writeBlock (tools/GecoIndex.java:119-146): member sequences in block order, each followed by '\\0'."""
    headers = [s.header for s in block.sequences]
    buf = np.zeros(sum(s.length + 1 for s in block.sequences), dtype=np.uint8)
    p = 0
    for s in block.sequences:
        seq = s.sequence() if callable(s.sequence) else s.sequence
        buf[p:p + s.length] = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray, memoryview)) else seq
        p += s.length + 1
    return headers, buf


def index_records(records: Iterable[tuple[str, object]], opath, xpath=None, sampling: int = 32,
                  devices: Sequence[int] = (0,)) -> dict:
    """GecoIndex.index on in-memory records [(header, bytes | uint8 array | callable returning one)]."""
    t1 = time.perf_counter()
    seqs = []
    for i, (h, s) in enumerate(records):
        length = s.length if hasattr(s, "length") else len(s)
        seqs.append(FastaSequence(h, length, s, i))
    blocks = merge_blocks(seqs)
    if not blocks:
        raise ValueError("no data found")
    with GecozFileWriter(opath, xpath, sampling, devices) as writer:
        for block in blocks:
            headers, text = block_text(block)
            writer.write(headers, text)
        timings = writer.timings
    return {"blocks": [[s.header for s in b.sequences] for b in blocks], "seconds": time.perf_counter() - t1,
            "timings": timings}


def index(ipath, opath, xpath=None, sampling: int = 32, devices: Sequence[int] = (0,)) -> dict:
    """GecoIndex.index(Path ipath, Path opath, Path xpath, int sampling, int threads)  :51-117."""
    return index_records(read_fasta(ipath), opath, xpath, sampling, devices)
