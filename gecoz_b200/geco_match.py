"""Query drivers: `gecotools -c / -s` for one pattern, and the batched GFF search for a file of patterns.

Mirrors of tools/GecoMatch.java:51-157 and tools/SimpleGFFGenerator.java:45-163
(/root/reference/java/nova-gecoz/src/main/java/es/elixir/bsc/ngs/nova/gecoz/tools/).  The reference prints to
stdout while it searches one pattern at a time; here the searches are batched per block on the GPU
(gcz_find_batch) and the same lines come back in the same order.
"""
from __future__ import annotations

from pathlib import Path
from typing import Iterable, Iterator, Sequence

import numpy as np

from .gecoz_file import GecozFileReader
from .gssa import pack_patterns


# ---- GecoMatch -----------------------------------------------------------------------------------------------------
def _print(headers: Sequence[str], res, match: bool, out: list[str]) -> int:
    """GecoMatch.print :143-157."""
    count = 0
    for i, r in enumerate(res):
        if r is not None and len(r) > 0:
            count += len(r)
            out.append(f">{headers[i]} found : {len(r)}")
            if match:
                out.extend(str(int(p)) for p in r)
    return count


def _match(ipath, header: str | None, pattern: str | bytes, match: bool, device: int = 0) -> tuple[list[str], int]:
    """GecoMatch.match(ipath, header, pattern, match) :59-141.  Returns (stdout lines, total found)."""
    ipath = Path(ipath)
    if not ipath.is_file():
        raise FileNotFoundError(f"no gecoz file found: {ipath}")
    if not GecozFileReader.checkFormat(ipath):
        raise ValueError(f"invalid gecoz file format: {ipath}")
    pat = pattern.encode("utf-8") if isinstance(pattern, str) else bytes(pattern)
    out: list[str] = []
    total = 0
    with GecozFileReader(ipath, device) as reader:
        if header is not None:                                            # :79-108
            bheader = reader.findBlockHeader(header)
            if bheader is None:
                raise KeyError(f"no sequence found: {header}")
            ssa = reader.read(bheader)
            try:
                res = ssa.find(pat)
            finally:
                ssa.close()
            if res is not None and len(res) > 0:
                nstr = bheader.findHeader(header)
                if res[nstr] is not None and len(res[nstr]) > 0:
                    total = len(res[nstr])
                    out.append(f">{header} found : {total}")
                    if match:
                        out.extend(str(int(p)) for p in res[nstr])
        else:                                                             # :109-134
            for bheader in reader.getBlockHeaders():
                ssa = reader.read(bheader)
                try:
                    res = ssa.find(pat)
                finally:
                    ssa.close()
                if res is not None and len(res) > 0:
                    total += _print(bheader.headers, res, match, out)
    return out, total


def match(ipath, header: str | None, pattern, device: int = 0) -> list[str]:
    """`gecotools -i x.gcz -s [header] PATTERN`: ">hdr found : k" then the 0-based positions (GecoMatch.match :51-53)."""
    return _match(ipath, header, pattern, True, device)[0]


def count(ipath, header: str | None, pattern, device: int = 0) -> list[str]:
    """`gecotools -i x.gcz -c [header] PATTERN` (GecoMatch.count :55-57)."""
    return _match(ipath, header, pattern, False, device)[0]


# ---- SimpleGFFGenerator ----------------------------------------------------------------------------------------------
_COMPLEMENT = np.arange(256, dtype=np.uint8)
for _a, _b in ((b"A", b"T"), (b"T", b"A"), (b"C", b"G"), (b"G", b"C")):
    _COMPLEMENT[_a[0]] = _b[0]


def read_patterns(path) -> Iterator[tuple[str, bytes]]:
    """The record loop of SimpleGFFGenerator.search :59-86: a line starting with '>' or '@' opens a record, one
    starting with '+' closes it (FASTQ qualities are skipped because no record is open), the other lines are
    appended verbatim; records without sequence bytes are dropped."""
    header, parts = None, []
    with open(path, "rb") as f:
        for raw in f.read().splitlines():                                 # BufferedReader.readLine: \\n, \\r or \\r\\n
            line = raw.decode("utf-8", errors="replace")
            if line.startswith(">") or line.startswith("@"):
                if header is not None and parts:
                    yield header, b"".join(parts)
                header, parts = line[1:], []
            elif line.startswith("+"):
                if header is not None and parts:
                    yield header, b"".join(parts)
                header, parts = None, []
            elif header is not None and raw:
                parts.append(raw)
    if header is not None and parts:
        yield header, b"".join(parts)


def _java_split_bar(header: str) -> list[str]:
    """String.split("\\\\|"): trailing empty strings are removed."""
    parts = header.split("|")
    while parts and parts[-1] == "":
        parts.pop()
    return parts if (parts or "|" in header) else [header]


def _attributes(header: str) -> str:
    h = _java_split_bar(header)
    s = ""
    if h:
        s += "ID=" + h[0]
    for note in h[1:]:
        s += ";Note=" + note
    return s


def gff_lines(gssas: Sequence, block_headers: Sequence[Sequence[str]], records: Iterable[tuple[str, bytes]]) -> list[str]:
    """SimpleGFFGenerator.search for in-memory records against open blocks: every record is searched as given
    (U -> T) and reverse-complemented, against every block; one GFF line per occurrence, in the reference's order
    (record, strand, block, string, position)."""
    recs = list(records)
    fwd, rev = [], []
    for _, seq in recs:
        a = np.frombuffer(bytes(seq), dtype=np.uint8).copy()
        a[a == ord("U")] = ord("T")                                        # :96-100
        fwd.append(a.tobytes())
        rev.append(_COMPLEMENT[a[::-1]].tobytes())                         # :104-108
    data, off = pack_patterns(fwd + rev)
    found = [g.find_batch_raw(packed=(data, off)) if g is not None else None for g in gssas]
    nrec = len(recs)
    lines: list[str] = []
    for r, (header, _) in enumerate(recs):
        attrs = _attributes(header)
        length = len(fwd[r])
        for strand, q in (("+", r), ("-", nrec + r)):
            for b, res in enumerate(found):
                if res is None:
                    continue
                per, pos, poff = res
                o = int(poff[q])
                for j in range(per.shape[1]):
                    k = int(per[q, j])
                    name = block_headers[b][j]
                    for p in pos[o:o + k]:
                        lines.append(f"{name}\tgecotools\tdna\t{int(p) + 1}\t{int(p) + length}\t1.000\t{strand}\t.\t{attrs}")
                    o += k
    return lines


def search(ref, fasta, device: int = 0) -> list[str]:
    """`gecotools -i x.gcz -s patterns.fa` (SimpleGFFGenerator.search(Path ref, Path fasta) :45-93)."""
    with GecozFileReader(ref, device) as reader:
        bheaders = reader.getBlockHeaders()
        gssas = [reader.read(h) for h in bheaders]
        try:
            return gff_lines(gssas, [h.headers for h in bheaders], read_patterns(fasta))
        finally:
            for g in gssas:
                if g is not None:
                    g.close()
