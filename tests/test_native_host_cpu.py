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


def test_gzipped_fasta_is_refused(tmp_path):
    import gzip
    p = tmp_path / "x.fa.gz"
    with gzip.open(p, "wb") as g:
        g.write(b">a\nACGT\n")
    with pytest.raises(N.GczFormatError):
        NF.Fasta(p)
    with NF.Fasta(gzip.open(p, "rb").read()) as f:                       # the caller decompresses
        assert f.records()[0][0] == "a"


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
        c = np.bincount(t, minlength=256).astype(np.int64)
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
