"""Extraction drivers: `.gcz` -> FASTA, or one subsequence to a raw file.

Mirror of tools/GecoRead.java:33-175 (/root/reference/java/nova-gecoz/.../tools/) and of the byte layout
fasta/FastaFileWriter.java:132-215 gives a multi-line sequence.  The reference extracts every sequence in calls of
4 MiB (SequenceExtractor :141-175); the same call boundaries are kept because each call restarts the LF walk at its
own sampled position — only then are the bytes the reference's in merged blocks as well.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from .gecoz_file import GecozFileReader

LINE_LENGTH = 50                    # fasta/FastaFileWriter.java:32
EXTRACT_BUFFER = 1024 * 1024 * 4    # tools/GecoRead.java:163


def fasta_record_bytes(header: str, seq: np.ndarray) -> bytes:
    """`>header\\n`, then what FastaSequenceWriter.run (:204-214) leaves in the slice reserved by
    FastaFileWriter.write(TFastaSequence) (:132-151, multiline): a line break after every 50 symbols and one more
    at the end (length + length / 50 + 1 bytes; the pipe reads are assumed to fill each line, which they do while
    the extractor is ahead of the writer)."""
    n = len(seq)
    body = np.full(n + n // LINE_LENGTH + 1, ord("\n"), dtype=np.uint8)
    idx = np.arange(n, dtype=np.int64)
    body[idx + idx // LINE_LENGTH] = seq
    return b">" + header.encode("utf-8") + b"\n" + body.tobytes()


def extract_sequence(ssa, nstr: int, length: int | None = None) -> np.ndarray:
    """SequenceExtractor.run :155-174: the whole string, one GSSA.extract call per 4 MiB."""
    length = ssa.getLength(nstr) if length is None else length
    out = np.zeros(max(length, 1), dtype=np.uint8)
    start = 0
    while True:                                                   # do { ... } while (from < len)
        got = ssa.extract(nstr, start, EXTRACT_BUFFER)
        out[start:start + len(got)] = got[:max(length - start, 0)]
        start += len(got)
        if start >= length or len(got) == 0:
            break
    return out[:length]


def fasta(ipath, opath, device: int = 0) -> int:
    """`gecotools -i x.gcz -o x.fa` (GecoRead.fasta :83-139): every sequence of every block, in file order."""
    ipath = Path(ipath)
    if not ipath.is_file():
        raise FileNotFoundError(f"no gecoz file found: {ipath}")
    if not GecozFileReader.checkFormat(ipath):
        raise ValueError(f"invalid gecoz file format: {ipath}")
    count = 0
    with GecozFileReader(ipath, device) as reader, open(opath, "wb") as out:
        for bheader in reader.getBlockHeaders():
            ssa = reader.read(bheader)
            try:
                for header in bheader.headers:
                    nstr = bheader.findHeader(header)            # the FIRST string with that header, like the reference
                    out.write(fasta_record_bytes(header, extract_sequence(ssa, nstr)))
                    count += 1
            finally:
                ssa.close()
    return count


def sequence(ipath, header: str, start: int, end: int, opath, device: int = 0) -> int:
    """`gecotools -i x.gcz -o out -s header from to` (GecoRead.sequence :33-81): the raw symbols [from, to) of one
    sequence, one GSSA.extract call."""
    with GecozFileReader(ipath, device) as reader:
        bheader = reader.findBlockHeader(header)
        if bheader is None:
            raise KeyError(f"no sequence found: {header}")
        ssa = reader.read(bheader)
        try:
            nstr = bheader.findHeader(header)
            end = min(end, ssa.getLength(nstr))
            data = ssa.extract(nstr, start, max(end - start, 0))
        finally:
            ssa.close()
    with open(opath, "wb") as f:
        f.write(data.tobytes())
    return len(data)
