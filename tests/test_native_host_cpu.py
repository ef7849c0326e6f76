"""The native host layer (include/gcz_file.h, csrc/host_file.cpp) on CPU: FASTA records against a literal Python
transliteration of FastaIterator, block planning against the oracle, the writer with the oracle injected as the
engine (whole-file parity without a GPU), the reader's block directory."""
import ctypes as C

import numpy as np
import pytest

from gecoz_b200 import _native as N
from gecoz_b200 import native_file as NF
from gecoz_b200 import synth
from gecoz_b200.gecoz_file import GecozRefBlockHeader, GecozSSABlockHeader
from gecoz_b200.geco_index import FastaSequence, merge_blocks
from oracle import gcz_oracle as O


def _fasta_iterator(buf: bytes):
    """fasta/FastaIterator.java:39-127 (lazy) + FastaFileReader.read :109-160, byte for byte, in Python."""
    p = 0

    def rd():
        nonlocal p
        if p < len(buf):
            p += 1
            return buf[p - 1]
        return -1

    ch, position, out = 13, 0, []
    while True:
        while ch >= 0 and ch not in (62, 64):
            ch = rd()
            position += 1
        if ch < 0:
            break
        header = bytearray()
        while True:
            ch = rd()
            if ch < 0 or ch == 10:
                break
            position += 1
            if ch != 13:
                header.append(ch)
        position += 1
        lines = length = 0
        posnew = position
        while True:
            if ch >= 0 and ch not in (13, 10):
                lines += 1
                while True:
                    posnew += 1
                    length += 1
                    ch = rd()
                    if ch < 0 or ch in (13, 10):
                        break
            posnew += 1
            ch = rd()
            if ch < 0 or ch in (62, 64, 43):
                break
        if ch == 43:
            qlines, qlength = -1, 0
            while True:
                while True:
                    ch = rd()
                    if ch < 0 or ch in (13, 10):
                        break
                    qlength += 1
                    posnew += 1
                posnew += 1
                qlines += 1
                if not (qlength < length and qlines < lines):
                    break
        if lines > 1:
            seq = bytes(c for c in buf[position:] if c not in (13, 10))[:length]
        else:
            seq = buf[position:position + length]
        out.append((header.decode("latin-1"), position, length, lines > 1, seq))
        position = posnew
    return out


FASTA_CASES = [
    b">a\nACGT\n>b desc\r\nAC\r\nGT\r\n\r\nTT\n",
    b"junk before\n>only header\n",
    b">x\nACGT",                                                        # no trailing newline
    b"@r1\nACGTN\n+\nIIIII\n@r2\nGG\nCC\n+r2\nII\nII\n>after\nTTTT\n",    # FASTQ, one- and two-line
    b"@r1\nACGT\n+r1 again\nIIII\n@r2\nTT\n+\n@I\n",                      # '+' line counted into the quality length
    b">e1\n>e2\n\n>s\nA\n",
    b"",
    b">h\n\n\nACGT\n\nAC\n",
]


@pytest.mark.parametrize("data", FASTA_CASES)
def test_fasta_records_follow_the_iterator(data, tmp_path):
    exp = _fasta_iterator(data)
    path = tmp_path / "x.fa"
    path.write_bytes(data)
    for src in (data, path):
        with NF.Fasta(src) as f:
            assert len(f) == len(exp)
            for i, (h, pos, ln, ml, seq) in enumerate(exp):
                assert f.record(i) == (h, pos, ln, ml)
                assert f.read(i).tobytes() == seq


def test_fasta_random_files():
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b"ACGTN>@+\r\n\n\nacgt ", np.uint8)
    for _ in range(200):
        data = alphabet[rng.integers(0, len(alphabet), int(rng.integers(0, 300)))].tobytes()
        exp = _fasta_iterator(data)
        with NF.Fasta(data) as f:
            assert [f.record(i) + (f.read(i).tobytes(),) for i in range(len(f))] == exp


def _scan(data: bytes):
    with NF.Fasta(data) as f:
        return [f.record(i) + (f.read(i).tobytes(),) for i in range(len(f))]


@pytest.mark.parametrize("slice_bytes,threads", [("5", "3"), ("64", "2"), ("1000000", "1")])
def test_fast_scanner_equals_the_literal_machine(monkeypatch, slice_bytes, threads):
    """scan_fasta (vector searches, sequence regions counted by several threads) against scan_fasta_literal (the Java loop,
    one character at a time) on inputs made of the characters the machine reacts to; tiny slices force the threaded steps."""
    rng = np.random.default_rng(int(slice_bytes) + int(threads))
    alphabets = [np.frombuffer(b"ACGTN>@+\r\n\n\nacgt ", np.uint8), np.frombuffer(b"AC\n\r>@+", np.uint8),
                 np.frombuffer(b"ACGTACGTACGTACGTACGTACGTACGTACGT\n>", np.uint8), np.frombuffer(b"A\n+@I", np.uint8)]
    for it in range(1200):
        alpha = alphabets[it % len(alphabets)]
        data = alpha[rng.integers(0, len(alpha), int(rng.integers(0, 2500 if it % 10 == 0 else 200)))].tobytes()
        monkeypatch.setenv("GCZ_FASTA_LITERAL", "1")
        exp = _scan(data)
        monkeypatch.delenv("GCZ_FASTA_LITERAL")
        monkeypatch.setenv("GCZ_FASTA_SLICE", slice_bytes)
        monkeypatch.setenv("GCZ_HOST_THREADS", threads)
        got = _scan(data)
        monkeypatch.delenv("GCZ_FASTA_SLICE")
        monkeypatch.delenv("GCZ_HOST_THREADS")
        assert got == exp, data


def test_long_multiline_sequence_is_assembled_by_several_threads(monkeypatch):
    """A record large enough for the threaded reader (>= 4 MiB of source per thread), mixed line ends and widths."""
    rng = np.random.default_rng(9)
    seq = synth.iid_acgtn(12_000_000, seed=3)
    parts, p = [b">big one\r\n"], 0
    while p < len(seq):
        w = int(rng.integers(1, 200)) if p < 50_000 else 61
        parts.append(seq[p:p + w].tobytes())
        parts.append((b"\n", b"\r\n", b"\n\n")[int(rng.integers(0, 3))] if p < 50_000 else b"\n")
        p += w
    parts.append(b">next\nAC\nGT\n")
    data = b"".join(parts)
    monkeypatch.setenv("GCZ_HOST_THREADS", "3")
    with NF.Fasta(data) as f:
        assert len(f) == 2 and f.record(0)[2:] == (len(seq), True)
        assert np.array_equal(f.read(0), seq)
        assert f.read(1).tobytes() == b"ACGT"
        short = np.zeros(len(seq) - 1, np.uint8)                       # a buffer shorter than the sequence is refused
        assert N.lib().gcz_fasta_read(f._h, 0, N.ptr(short), len(short)) == N.GCZ_E_ARG


def test_gzipped_fasta(tmp_path):
    """FastaFileReader probes for GZIP and reads the decompressed stream (fasta/FastaFileReader.java:71-96); here through
    zlib looked up at run time, multi-member files (BGZF) included."""
    import gzip
    data = b">a\nACGT\nAC\n>b desc\nGG\n"
    p1, p2 = tmp_path / "x.fa.gz", tmp_path / "y.fa.gz"
    with gzip.open(p1, "wb") as g:
        g.write(data)
    p2.write_bytes(gzip.compress(data[:11]) + gzip.compress(data[11:]))
    for p in (p1, p2):
        with NF.Fasta(p) as f:
            assert [(h, s.tobytes()) for h, s in f.records()] == [("a", b"ACGTAC"), ("b desc", b"GG")]
    bad = tmp_path / "bad.fa.gz"
    bad.write_bytes(gzip.compress(data)[:-6])
    with pytest.raises(N.GczFormatError):
        NF.Fasta(bad)


def test_plan_blocks_matches_oracle_and_python():
    rng = np.random.default_rng(9)
    cases = [(synth.HG38_LENGTHS, synth.HG38_NAMES)]
    for _ in range(40):
        n = int(rng.integers(1, 30))
        lengths = [int(x) for x in rng.integers(0, 2000, n)]
        headers = [f"s{int(x)}" for x in rng.integers(0, max(2, n // 2), n)]      # duplicate (length, header) pairs happen
        cases.append((lengths, headers))
    for lengths, headers in cases:
        got = NF.plan_blocks(lengths, headers)
        py = merge_blocks([FastaSequence(h, ln, None, i) for i, (h, ln) in enumerate(zip(headers, lengths))])
        assert [[(s.length, s.header) for s in b.sequences] for b in py] == [[(lengths[i], headers[i]) for i in b] for b in got]
        if len(set(zip(lengths, headers))) == len(lengths):                      # the oracle's ids are ambiguous otherwise
            assert O.merge_blocks(lengths, headers) == got
    assert len(NF.plan_blocks(synth.HG38_LENGTHS, synth.HG38_NAMES)) == 18


def test_headers_match_python_mirror():
    for headers in (["chr1"], ["a b c", "x", ""], ["h" * 300, "chrM"]):
        assert NF.ref_header(headers, 123456789012, 987654321) == GecozRefBlockHeader(headers, 123456789012, 987654321).to_bytes()
        assert NF.ssa_header(headers, 55555) == GecozSSABlockHeader(headers, 55555).to_bytes()
        assert NF.header_hash(headers) == GecozRefBlockHeader.block_header_hash(headers)


def _oracle_engine(fail_first_build_with=None):
    state = {"builds": 0}

    def count(device, text, n, counts):
        t = np.ctypeslib.as_array(C.cast(text, C.POINTER(C.c_uint8)), shape=(n,))
        c = np.bincount(t, minlength=256).astype(np.int64)                                    # kept alive across the copy
        C.memmove(counts, c.ctypes.data, 256 * 8)
        return 0

    def build(device, text, n, rate, shape, gcz, gcz_len, gcx, gcx_len, sa, bwt):
        state["builds"] += 1
        if fail_first_build_with is not None and state["builds"] == 1:
            return fail_first_build_with
        t = np.ctypeslib.as_array(C.cast(text, C.POINTER(C.c_uint8)), shape=(n,)).copy()
        r = O.build_block(t, rate)
        assert len(r["gcz_body"]) == gcz_len and len(r["gcx_body"]) == gcx_len
        C.memmove(gcz, r["gcz_body"].ctypes.data, gcz_len)
        C.memmove(gcx, r["gcx_body"].ctypes.data, gcx_len)
        return 0

    eng = N.Engine(N.COUNT_SYMBOLS_FN(count), N.BUILD_BLOCK_FN(build))
    return eng, state


def _write_fasta(path, recs, width=60, eol=b"\n"):
    with open(path, "wb") as f:
        for h, s in recs:
            f.write(b">" + h.encode() + eol)
            for i in range(0, len(s), width):
                f.write(s[i:i + width].tobytes() + eol)


@pytest.mark.parametrize("rate,devices", [(32, (0,)), (4, (0, 1, 2))])
def test_native_writer_with_the_oracle_engine(tmp_path, rate, devices):
    """gcz_index_fasta end to end without a GPU: FASTA -> records -> blocks -> offsets / headers / mapped slices ->
    engine.  The files must be the oracle's, whatever the number of blocks in flight."""
    recs = [(f"seq{i} some description", synth.iid_acgtn(int(ln), 40 + i)) for i, ln in
            enumerate([9_000, 6_100, 3_050, 3_000, 900, 500, 20, 20, 0])]
    fa = tmp_path / "x.fa"
    _write_fasta(fa, recs, eol=b"\r\n")
    eng, state = _oracle_engine()
    with NF.Fasta(fa) as f:
        rep = NF.index(f, tmp_path / "x.gcz", sampling=rate, devices=devices, engine=eng)
    kept = [(h, s.tobytes()) for h, s in recs]
    gcz, gcx, blocks = O.write_files(kept, rate)
    assert (tmp_path / "x.gcz").read_bytes() == gcz
    assert (tmp_path / "x.gcx").read_bytes() == gcx
    assert rep["blocks"] == len(blocks) == state["builds"] and rep["sequences"] == len(recs)
    assert rep["symbols"] == sum(len(s) + 1 for _, s in recs)

    with NF.Reader(tmp_path / "x.gcz") as r:                               # and the native reader walks them
        assert r.n_blocks == len(blocks)
        assert r.sampling_factor == rate.bit_length() - 1
        for b, ids in enumerate(blocks):
            info = r.block(b)
            assert info["headers"] == [recs[i][0] for i in ids]
            assert info["len"] == sum(len(recs[i][1]) + 1 for i in ids)
        b, s = r.find(recs[4][0])
        assert blocks[b][s] == 4
        with pytest.raises(N.GczError):
            r.find("no such header")


def test_native_writer_many_blocks_and_a_large_alphabet(tmp_path):
    """More blocks than body buffers (workers wait for one to come back) and bodies larger than the pool's DNA estimate
    (the prefilled buffers are dropped for larger ones)."""
    rng = np.random.default_rng(3)
    alpha = np.frombuffer(b"ACGTNacgtnRYKMSWBDHVrykmswbdhv", np.uint8)
    recs = [(f"s{i}", alpha[rng.integers(0, len(alpha), 2000 - i)]) for i in range(30)]
    fa = tmp_path / "m.fa"
    _write_fasta(fa, recs, width=70)
    eng, state = _oracle_engine()
    with NF.Fasta(fa) as f:
        rep = NF.index(f, tmp_path / "m.gcz", sampling=8, devices=(0,), engine=eng)
    gcz, gcx, blocks = O.write_files([(h, s.tobytes()) for h, s in recs], 8)
    assert len(blocks) == 30 == rep["blocks"] == state["builds"]
    assert (tmp_path / "m.gcz").read_bytes() == gcz and (tmp_path / "m.gcx").read_bytes() == gcx


def test_native_writer_with_a_concurrent_engine(tmp_path):
    """The same pipeline with an engine that does not hold the GIL (tools/hostbench.py's stand-in, compiled with gcc):
    builds, buffer hand-overs and file writes really overlap.  Every block's header and both bodies must be where the
    reader expects them."""
    import sys
    sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent / "tools"))
    import hostbench
    eng = hostbench.standin_engine(50.0)
    recs = [(f"c{i}", synth.iid_acgtn(40_000 - 137 * i, 70 + i)) for i in range(40)]
    fa = tmp_path / "c.fa"
    _write_fasta(fa, recs)
    for devices in ((0,), (0, 1, 2, 3)):
        with NF.Fasta(fa) as f:
            rep = NF.index(f, tmp_path / "c.gcz", sampling=32, devices=devices, engine=eng)
        assert rep["blocks"] == 40
        ref, ssa = (tmp_path / "c.gcz").read_bytes(), (tmp_path / "c.gcx").read_bytes()
        rp = sp = 0
        for b in range(40):
            n = 40_000 - 137 * b + 1                                       # file order: longest first
            assert ref[rp:rp + 8] == b"GecozBWT" and ssa[sp:sp + 8] == b"GecozSSA"
            size, text_len = int.from_bytes(ref[rp + 9:rp + 17], "little"), int.from_bytes(ref[rp + 17:rp + 25], "little")
            assert text_len == n
            hlen = GecozRefBlockHeader.block_header_length([f"c{b}"])
            assert set(ref[rp + hlen:rp + size]) == {n % 251}
            idx = int.from_bytes(ssa[sp + 9:sp + 17], "little")
            assert idx == O.index_size(n, 5) and set(ssa[sp + 25:sp + 25 + idx]) == {n % 241}
            rp, sp = rp + size, sp + 25 + idx
        assert rp == len(ref) and sp == len(ssa)


def test_native_writer_retries_a_block_that_ran_out_of_memory(tmp_path):
    recs = [("a", synth.iid_acgtn(3000, 1)), ("b", synth.iid_acgtn(2900, 2))]
    fa = tmp_path / "y.fa"
    _write_fasta(fa, recs)
    eng, state = _oracle_engine(fail_first_build_with=N.GCZ_E_NOMEM)
    with NF.Fasta(fa) as f:
        NF.index(f, tmp_path / "y.gcz", engine=eng)
    gcz, gcx, blocks = O.write_files([(h, s.tobytes()) for h, s in recs])
    assert (tmp_path / "y.gcz").read_bytes() == gcz and (tmp_path / "y.gcx").read_bytes() == gcx
    assert state["builds"] == len(blocks) + 1
    eng, _ = _oracle_engine(fail_first_build_with=N.GCZ_E_CUDA)            # any other error surfaces
    with NF.Fasta(fa) as f, pytest.raises(N.GczError):
        NF.index(f, tmp_path / "z.gcz", engine=eng)


def test_native_writer_empty_input(tmp_path):
    with NF.Fasta(b"no records here\n") as f, pytest.raises(N.GczError, match="no data found"):
        NF.index(f, tmp_path / "e.gcz", engine=_oracle_engine()[0])


# ---- the callers (gcz_match, gcz_gff_search, gcz_extract_fasta) with the oracle as the query engine ---------------------
class _OracleQueryEngine:
    """gcz_query_engine whose members call the oracle: lets the native callers run without a GPU."""

    def __init__(self):
        self.libc = C.CDLL(None)
        self.libc.malloc.restype = C.c_void_p
        self.libc.malloc.argtypes = [C.c_size_t]
        self.libc.free.argtypes = [C.c_void_p]
        self.open = {}
        self.next_id = 1
        E = N.QueryEngine
        types = dict(E._fields_)
        self.struct = E(types["open_block"](self._open), types["close_block"](self._close), types["num_strings"](self._num_strings),
                        types["string_ends"](self._string_ends), types["find_batch"](self._find), types["extract"](self._extract),
                        types["release"](self._release))

    def _open(self, device, gcz, gcz_len, text_len, gcx, gcx_len, out):
        a = np.ctypeslib.as_array(C.cast(gcz, C.POINTER(C.c_uint8)), shape=(gcz_len,)).copy()
        b = np.ctypeslib.as_array(C.cast(gcx, C.POINTER(C.c_uint8)), shape=(gcx_len,)).copy()
        self.open[self.next_id] = O.GSSA(a, text_len, b)
        out[0] = self.next_id
        self.next_id += 1
        return 0

    def _close(self, h):
        self.open.pop(h).close()

    def _num_strings(self, h, out):
        out[0] = self.open[h].n_strings
        return 0

    def _string_ends(self, h, e):
        ends = self.open[h].string_ends()
        C.memmove(e, ends.ctypes.data, ends.nbytes)
        return 0

    def _find(self, h, pats, off, n, per, pos_out, off_out):
        g = self.open[h]
        o = np.ctypeslib.as_array(off, shape=(n + 1,))
        data = np.ctypeslib.as_array(C.cast(pats, C.POINTER(C.c_uint8)), shape=(max(int(o[n]), 1),))
        counts, positions, offs = [], [], [0]
        for i in range(n):
            res = g.find(data[o[i]:o[i + 1]].tobytes())
            row = [0] * g.n_strings
            if res is not None:
                for s, a in enumerate(res):
                    if a is not None:
                        row[s] = len(a)
                        positions.extend(int(x) for x in a)
            counts.extend(row)
            offs.append(len(positions))
        for k, v in enumerate(counts):
            per[k] = v
        pbuf = self.libc.malloc(8 * max(len(positions), 1))
        obuf = self.libc.malloc(8 * (n + 1))
        parr, oarr = np.asarray(positions + [0], np.int64), np.asarray(offs, np.int64)      # kept alive across the copies
        C.memmove(pbuf, parr.ctypes.data, 8 * max(len(positions), 1))
        C.memmove(obuf, oarr.ctypes.data, 8 * (n + 1))
        pos_out[0], off_out[0] = pbuf, obuf
        return 0

    def _extract(self, h, nstr, start, out, cap, written):
        try:
            got = self.open[h].extract(nstr, start, cap)
        except (ValueError, IndexError):
            return N.GCZ_E_RANGE
        C.memmove(out, got.ctypes.data, len(got))
        written[0] = len(got)
        return 0

    def _release(self, p):
        self.libc.free(p)


@pytest.fixture(scope="module")
def small_index(tmp_path_factory):
    d = tmp_path_factory.mktemp("native_callers")
    recs = [(f"chr{i} test", synth.iid_acgtn(int(ln), 60 + i, p_n=0.0)) for i, ln in enumerate([4_000, 2_500, 1_200, 1_100, 150, 100])]
    _write_fasta(d / "g.fa", recs)
    eng, _ = _oracle_engine()
    with NF.Fasta(d / "g.fa") as f:
        NF.index(f, d / "g.gcz", engine=eng)
    _, _, blocks = O.write_files([(h, s.tobytes()) for h, s in recs])
    oracle_blocks = []
    for ids in blocks:
        text = synth.block_of([recs[i][1] for i in ids])
        r = O.build_block(text, 32)
        oracle_blocks.append(([recs[i][0] for i in ids], O.GSSA(r["gcz_body"], len(text), r["gcx_body"])))
    return d, recs, oracle_blocks


def test_native_match_lines(small_index):
    d, recs, blocks = small_index
    q = _OracleQueryEngine()
    pats = [recs[0][1][100:108].tobytes(), recs[3][1][5:11].tobytes(), recs[5][1][:4].tobytes(), b"ACGTACGTACGTACGTACGTAC", b"A"]
    with NF.Reader(d / "g.gcz") as r:
        for pat in pats:
            for want_pos in (False, True):
                exp = []
                for headers, og in blocks:
                    res = og.find(pat)
                    for h, a in zip(headers, res or []):
                        if a is not None and len(a):
                            exp.append(f">{h} found : {len(a)}")
                            if want_pos:
                                exp += [str(int(p)) for p in a]
                assert r.match(None, pat, want_pos, engine=q.struct) == exp
        hdr = recs[3][0]
        headers, og = next(b for b in blocks if hdr in b[0])
        res = og.find(pats[1])
        a = res[headers.index(hdr)] if res is not None else None
        exp = ([f">{hdr} found : {len(a)}"] + [str(int(p)) for p in a]) if a is not None and len(a) else []
        assert r.match(hdr, pats[1], True, engine=q.struct) == exp
        with pytest.raises(N.GczError):
            r.match("chrZ", b"ACGT", engine=q.struct)
    assert not q.open                                                       # every block was closed again


def test_native_gff_lines(small_index):
    from gecoz_b200.geco_match import _attributes
    d, recs, blocks = small_index
    q = _OracleQueryEngine()
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    s0, s2 = recs[0][1], recs[2][1]
    pats = [("p1|first note|second", s0[200:230].tobytes()), ("rna", s0[300:320].tobytes().replace(b"T", b"U")),
            ("revcomp|", s2[50:75].tobytes().translate(comp)[::-1]), ("absent", b"ACGT" * 9), ("|", s2[10:22].tobytes()), ("short", b"ACG")]
    data = b""
    for i, (h, s) in enumerate(pats):
        data += (b"@" + h.encode() + b"\r\n" + s[:10] + b"\r\n" + s[10:] + b"\n+\n" + b"I" * len(s) + b"\n") if i == 1 else (b">" + h.encode() + b"\n" + s + b"\n")
    data += b">empty\n>last\nACGTAC"
    exp = []
    for h, s in pats + [("last", b"ACGTAC")]:
        fwd = s.replace(b"U", b"T")
        for strand, seq in (("+", fwd), ("-", fwd.translate(comp)[::-1])):
            for headers, og in blocks:
                res = og.find(seq)
                for name, a in zip(headers, res or []):
                    if a is not None:
                        exp += [f"{name}\tgecotools\tdna\t{int(p) + 1}\t{int(p) + len(seq)}\t1.000\t{strand}\t.\t{_attributes(h)}" for p in a]
    with NF.Reader(d / "g.gcz") as r:
        got = r.gff_search(data, engine=q.struct)
        assert r.gff_search(b"", engine=q.struct) == []
    assert got == exp and len(got) > 10


def test_native_extract_fasta(small_index, tmp_path):
    d, recs, blocks = small_index
    q = _OracleQueryEngine()
    with NF.Reader(d / "g.gcz") as r:
        assert r.extract_fasta(tmp_path / "out.fa", engine=q.struct) == len(recs)
    from gecoz_b200.geco_read import fasta_record_bytes
    exp = b""
    for headers, og in blocks:
        ends = og.string_ends()
        for nstr, h in enumerate(headers):
            length = int(ends[nstr] - (ends[nstr - 1] + 1 if nstr else 0))
            exp += fasta_record_bytes(h, og.extract(nstr, 0, 4 * 1024 * 1024)[:length])
    assert (tmp_path / "out.fa").read_bytes() == exp
    by_header = {h: s for h, s in recs}
    with NF.Fasta(tmp_path / "out.fa") as f:                               # single-string blocks come back verbatim
        back = dict(f.records())
    first = blocks[0][0][0]
    assert np.array_equal(back[first], by_header[first])


def test_native_extract_sequence(small_index, tmp_path):
    """gcz_extract_sequence == GecoRead.sequence: [from, min(to, length)) of one sequence with one extract call."""
    d, recs, blocks = small_index
    q = _OracleQueryEngine()
    out = tmp_path / "s.seq"
    with NF.Reader(d / "g.gcz") as r:
        for headers, og in blocks:
            ends = og.string_ends()
            for nstr, h in enumerate(headers):
                length = int(ends[nstr] - (ends[nstr - 1] + 1 if nstr else 0))
                for start, end in ((0, 2 ** 31 - 1), (7, 61), (length - 3, length + 10), (5, 5)):
                    if start < 0 or start > length:
                        continue
                    n = r.extract_sequence(h, start, end, out, engine=q.struct)
                    want = og.extract(nstr, start, max(min(end, length) - start, 0)) if min(end, length) > start else np.zeros(0, np.uint8)
                    assert n == len(want) and out.read_bytes() == want.tobytes()
        first = blocks[0][0][0]                                            # a single-string block: the text itself
        r.extract_sequence(first, 10, 90, out, engine=q.struct)
        assert out.read_bytes() == dict(recs)[first][10:90].tobytes()
        with pytest.raises(N.GczError):
            r.extract_sequence("no such header", 0, 10, out, engine=q.struct)
        with pytest.raises(N.GczError):
            r.extract_sequence(first, 50, 40, out, engine=q.struct)
