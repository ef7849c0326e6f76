"""GPU parity tests: every stage of the CUDA path, called through the C ABI, against the CPU oracle.

Bit-exact everywhere (integer / byte / index work).  Sizes are what the oracle finishes in seconds; the
BASELINE.json full sizes are covered by size-independent properties in test_gpu_fullsize.py.
"""
import ctypes as C
import json
import os
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = Path(__file__).parent / "golden"


@pytest.fixture(scope="module")
def G():
    import gecoz_b200 as g
    g.lib()                      # fails loudly when libgcz_b200.so is missing
    return g


@pytest.fixture(scope="module")
def O():
    from oracle import gcz_oracle as o
    o.lib()
    return o


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---- radix sort -------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,bits", [(1, 64), (7, 64), (4095, 64), (4096, 64), (4097, 64), (6144, 64), (100_000, 64),
                                    (1_000_003, 64), (300_000, 41), (250_000, 8), (3_000_000, 63)])
def test_sort_pairs(G, n, bits):
    rng = np.random.default_rng(n + bits)
    keys = rng.integers(0, 2 ** 63, n, dtype=np.uint64)
    if bits < 64:
        keys &= np.uint64((1 << bits) - 1)
    if n > 1000:
        keys[: n // 3] = keys[n // 3: 2 * (n // 3)]       # plenty of duplicates: stability matters
    vals = np.arange(n, dtype=np.uint32)
    k2, v2 = keys.copy(), vals.copy()
    G._native.check(G.lib().gcz_dbg_sort_pairs(0, _p(k2), _p(v2), n, 0, bits))
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k2, keys[order])
    assert np.array_equal(v2, vals[order])


def test_sort_skewed_digits(G):
    n = 500_000
    keys = np.zeros(n, dtype=np.uint64)
    keys[::7] = 1 << 40
    keys[::1001] = (1 << 62) + 5
    vals = np.arange(n, dtype=np.uint32)
    k2, v2 = keys.copy(), vals.copy()
    G._native.check(G.lib().gcz_dbg_sort_pairs(0, _p(k2), _p(v2), n, 0, 64))
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k2, keys[order]) and np.array_equal(v2, vals[order])


@pytest.mark.parametrize("n,bits", [(5, 64), (4097, 64), (1_000_003, 64), (3_000_000, 24)])
def test_sort_pairs_wide_status(G, monkeypatch, n, bits):
    """Sorts of 2^30 pairs and more use 64-bit status words in the look-back (radix_sort.cu: Status<unsigned long long>); the switch
    sends a small sort through that instantiation."""
    monkeypatch.setenv("GCZ_SORT_WIDE_STATUS", "1")
    test_sort_pairs(G, n, bits)
    monkeypatch.delenv("GCZ_SORT_WIDE_STATUS")
    test_sort_pairs(G, n, bits)


# ---- ranked bit vector layout ---------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 7, 64, 511, 512, 513, 65535, 65536, 65537, 70001, 131072, 1_000_001])
def test_ranked_vector(G, O, n):
    rng = np.random.default_rng(n)
    bits = (rng.random(n) < 0.4).astype(np.uint8)
    out = np.zeros(O.ranked_bytes(n) + 16, np.uint8)
    out[-16:] = 0xEE
    G._native.check(G.lib().gcz_dbg_ranked_vector(0, _p(bits), n, _p(out)))
    exp = O.ranked_write(bits)
    assert len(exp) == G.lib().gcz_ranked_bytes(n)
    assert np.array_equal(out[:len(exp)], exp)
    assert (out[-16:] == 0xEE).all()


# ---- IndexWaveletTree ---------------------------------------------------------------------------------
def test_iwt_pdf_table3(G):
    k = json.loads((GOLD / "reference_kats.json").read_text())["iwt_table3"]
    vals = np.array(k["values"], dtype=np.int32)
    out = np.zeros(15, np.uint8)
    G._native.check(G.lib().gcz_dbg_index_wavelet_tree(0, _p(vals), len(vals), _p(out)))
    assert out.tobytes().hex() == k["serialized_hex"]


@pytest.mark.parametrize("m", [1, 2, 3, 31, 32, 33, 1000, 65536, 65537, 500_001])
def test_iwt_random(G, O, m):
    vals = np.random.default_rng(m).permutation(m).astype(np.int32)
    exp = O.iwt_write(vals)
    out = np.zeros(len(exp), np.uint8)
    G._native.check(G.lib().gcz_dbg_index_wavelet_tree(0, _p(vals), m, _p(out)))
    assert np.array_equal(out, exp)


# ---- suffix array ----------------------------------------------------------------------------------------
def _texts():
    from gecoz_b200 import synth
    rng = np.random.default_rng(11)
    yield "tiny", np.frombuffer(b"GATTACA\0", np.uint8).copy()
    yield "one", np.frombuffer(b"\0", np.uint8).copy()
    yield "two_strings", np.frombuffer(b"ACGTN\0ACG\0", np.uint8).copy()
    yield "iid_100k", synth.cfg1_text(100_000)
    yield "chr_shaped_2M", synth.cfg2_text(2_000_000)                    # long N runs: deep prefix doubling
    yield "all_same", synth.block_of([np.full(50_000, ord("A"), np.uint8)])
    yield "tandem", synth.block_of([np.frombuffer(b"ACACACACGT" * 30_000, np.uint8)])
    yield "multi", synth.block_of([synth.iid_acgtn(30_000, 5), synth.iid_acgtn(20_000, 6), synth.iid_acgtn(7, 7),
                                   np.zeros(0, np.uint8), synth.iid_acgtn(20_000, 6)])   # empty + duplicate sequences
    yield "bytes", np.concatenate([rng.integers(1, 255, 80_000, dtype=np.uint8), np.zeros(1, np.uint8)])
    yield "lower_iupac", synth.block_of([np.frombuffer(b"ACGTNacgtnRYKM", np.uint8)[rng.integers(0, 14, 150_000)]])
    # long runs of one symbol take the closed-form path of the sorter (suffix_sort.cu, point 5)
    acgtn = np.frombuffer(b"ACGTN", np.uint8)
    lens = rng.integers(1, 200, 4000)
    yield "runs_mixed", synth.block_of([np.repeat(acgtn[rng.integers(0, 5, len(lens))], lens)])
    parts = []
    for i in range(600):                               # equal-length N runs: ties on the run length, both sides
        parts += [np.full(60 + (i % 3), ord("N"), np.uint8), acgtn[rng.integers(0, 4, 1 + i % 7)]]
    yield "runs_equal", synth.block_of([np.concatenate(parts), np.full(64, ord("T"), np.uint8), np.full(64, ord("A"), np.uint8)])
    yield "runs_at_ends", synth.block_of([np.concatenate([np.full(5000, ord("N"), np.uint8), synth.iid_acgtn(20_000, 3),
                                                          np.full(7000, ord("N"), np.uint8)])])
    yield "separator_runs", synth.block_of([np.zeros(0, np.uint8)] * 40 + [synth.iid_acgtn(3000, 8)] + [np.zeros(0, np.uint8)] * 50)
    lens = rng.integers(3, 6, 30_000)                  # more long runs than the mark buffer holds: plain doubling
    yield "runs_bytes_overflow", np.concatenate([np.repeat(rng.integers(1, 255, len(lens), dtype=np.uint8), lens), np.zeros(1, np.uint8)])
    yield "periodic", synth.block_of([np.frombuffer(b"ACGT" * 40_000, np.uint8)])
    # > 32 internal tree nodes: the general node-bit emitter
    sym = np.arange(48, 48 + 44, dtype=np.uint8)
    yield "ascii44", synth.block_of([sym[np.minimum(rng.geometric(0.12, 120_000) - 1, 43)]])


@pytest.mark.parametrize("name,text", list(_texts()), ids=[t[0] for t in _texts()])
def test_suffix_array(G, O, name, text):
    sa = np.zeros(len(text), np.int32)
    G._native.check(G.lib().gcz_dbg_suffix_array(0, _p(text), len(text), _p(sa)))
    assert np.array_equal(sa, O.suffix_array(text))


def test_suffix_array_with_counter_flushes_and_wide_status(G, O, monkeypatch):
    """Two paths of the first sort that only very large blocks reach: the 16-bit lane counters of text_hist_kernel are summed
    into the histogram in mid-run (here after every tile instead of every 500), and the look-back runs on 64-bit status words."""
    from gecoz_b200 import synth
    text = synth.cfg2_text(3_000_000, seed=4)
    exp = O.suffix_array(text)
    for env in ("GCZ_TEXT_FLUSH_TILES", "GCZ_SORT_WIDE_STATUS"):
        monkeypatch.setenv(env, "1")
        sa = np.zeros(len(text), np.int32)
        G._native.check(G.lib().gcz_dbg_suffix_array(0, _p(text), len(text), _p(sa)))
        monkeypatch.delenv(env)
        assert np.array_equal(sa, exp), env


# ---- whole block: SA, BWT, .gcz body, .gcx body ------------------------------------------------------------
def _build(G, text, rate=32, want=True):
    counts = G.symbol_counts(text)
    assert np.array_equal(counts, np.bincount(text, minlength=256))
    shape = G.shape_from_counts(counts)
    gcz = np.zeros(shape.size + 32, np.uint8)
    gcx = np.zeros(G.index_size(len(text), rate.bit_length() - 1) + 32, np.uint8)
    gcz[-32:] = 0xEE
    gcx[-32:] = 0xEE
    sa = np.zeros(len(text), np.int32) if want else None
    bwt = np.zeros(len(text), np.uint8) if want else None
    t = G.build_block(0, text, len(text), rate, shape, gcz, gcx, sa, bwt)
    assert (gcz[-32:] == 0xEE).all() and (gcx[-32:] == 0xEE).all()        # nothing written past the slices
    return shape, gcz[:-32], gcx[:-32], sa, bwt, t


@pytest.mark.parametrize("name,text", list(_texts()), ids=[t[0] for t in _texts()])
@pytest.mark.parametrize("rate", [32, 4])
def test_build_block(G, O, name, text, rate):
    if name in ("bytes", "runs_bytes_overflow"):
        pytest.skip("alphabets whose length table needs a >7-bit code-length code are refused like the reference")
    shape, gcz, gcx, sa, bwt, t = _build(G, text, rate)
    ref = O.build_block(text, rate, want_sa=True, want_bwt=True)
    assert np.array_equal(sa, ref["sa"])
    assert np.array_equal(bwt, ref["bwt"])
    assert len(gcz) == len(ref["gcz_body"]) and np.array_equal(gcz, ref["gcz_body"])
    assert len(gcx) == len(ref["gcx_body"]) and np.array_equal(gcx, ref["gcx_body"])
    assert t["kernel_launches"] > 0


def test_build_block_gather_path(G, O, monkeypatch):
    """Blocks whose positions leave no room for the carried BWT symbol gather it from the text instead."""
    from gecoz_b200 import synth
    monkeypatch.setenv("GCZ_BWT_GATHER", "1")
    text = synth.cfg2_text(400_000, seed=5)
    shape, gcz, gcx, sa, bwt, t = _build(G, text, 32)
    ref = O.build_block(text, 32, want_sa=True, want_bwt=True)
    assert np.array_equal(sa, ref["sa"]) and np.array_equal(bwt, ref["bwt"])
    assert np.array_equal(gcz, ref["gcz_body"]) and np.array_equal(gcx, ref["gcx_body"])
    # and without the parity artefacts (the carried entries are then never cleaned)
    monkeypatch.delenv("GCZ_BWT_GATHER")
    shape, gcz, gcx, _, _, _ = _build(G, text, 32, want=False)
    assert np.array_equal(gcz, ref["gcz_body"]) and np.array_equal(gcx, ref["gcx_body"])


def test_text_staging_follows_the_buffer(G, O):
    """gcz_count_symbols leaves the upload for the build of the SAME host buffer; any other buffer is uploaded."""
    from gecoz_b200 import synth
    a = synth.cfg2_text(200_000, seed=21)
    b = synth.cfg2_text(200_000, seed=22)
    assert len(a) == len(b)
    shape_b = G.shape_from_counts(np.bincount(b, minlength=256).astype(np.int64))
    G.symbol_counts(a)                                       # stages a
    gcz = np.zeros(shape_b.size, np.uint8)
    gcx = np.zeros(G.index_size(len(b), 5), np.uint8)
    G.build_block(0, b, len(b), 32, shape_b, gcz, gcx)       # same length, other buffer: must not see a
    ref = O.build_block(b, 32)
    assert np.array_equal(gcz, ref["gcz_body"]) and np.array_equal(gcx, ref["gcx_body"])
    shape_a = G.shape_from_counts(G.symbol_counts(a))        # staged and used
    gcz = np.zeros(shape_a.size, np.uint8)
    G.build_block(0, a, len(a), 32, shape_a, gcz, gcx)
    ref = O.build_block(a, 32)
    assert np.array_equal(gcz, ref["gcz_body"]) and np.array_equal(gcx, ref["gcx_body"])


def test_stale_staged_text_is_not_used(G, O):
    """A buffer that changed after gcz_count_symbols (or was reallocated at the same address) is uploaded again."""
    from gecoz_b200 import synth
    a = synth.cfg2_text(150_000, seed=31)
    G.symbol_counts(a)                                       # stages the old content under a's address
    a[:-1] = synth.cfg2_text(150_000, seed=32)[:-1]          # same buffer, new text
    shape = G.shape_from_counts(np.bincount(a, minlength=256).astype(np.int64))
    gcz = np.zeros(shape.size, np.uint8)
    gcx = np.zeros(G.index_size(len(a), 5), np.uint8)
    G.build_block(0, a, len(a), 32, shape, gcz, gcx)
    ref = O.build_block(a, 32)
    assert np.array_equal(gcz, ref["gcz_body"]) and np.array_equal(gcx, ref["gcx_body"])


def test_provisional_golden_blocks(G):
    prov = json.loads((GOLD / "provisional_blocks.json").read_text())
    for name, p in prov.items():
        if name.startswith("_"):
            continue
        text = np.frombuffer(b"".join(s.encode() + b"\0" for s in p["sequences"]), np.uint8).copy()
        shape, gcz, gcx, sa, bwt, _ = _build(G, text, p["sampling_rate"])
        hdr = G.GecozRefBlockHeader(p["headers"], G.GecozRefBlockHeader.block_header_length(p["headers"]) + shape.size, len(text))
        assert (hdr.to_bytes() + gcz.tobytes()).hex() == p["gcz_hex"]
        assert (G.GecozSSABlockHeader(p["headers"], len(gcx)).to_bytes() + gcx.tobytes()).hex() == p["gcx_hex"]
        assert sa.tolist() == p["sa"] and bwt.tobytes().hex() == p["bwt_hex"]


# ---- queries ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def block_1m(G, O):
    from gecoz_b200 import synth
    text = synth.cfg2_text(1_000_000, seed=21)
    ref = O.build_block(text, 32, want_sa=True)
    return text, ref


def test_open_and_tables(G, O, block_1m):
    text, ref = block_1m
    og = O.GSSA(ref["gcz_body"], len(text), ref["gcx_body"])
    g = G.GSSA.open(0, ref["gcz_body"], len(text), ref["gcx_body"])
    assert g.length == len(text) and g.sampling_factor == og.sampling_factor == 5
    assert np.array_equal(g.c, og.c_array())
    assert g.e.tolist() == og.string_ends().tolist() == [len(text) - 1]
    g.close()


@pytest.mark.parametrize("table", [True, False], ids=["kmer_table", "step_by_step"])
def test_count_batch(G, O, block_1m, monkeypatch, table):
    """(sp, ep) of every pattern as GSSA.search leaves them, found or not — with the interval table of the block's 6-symbol strings
    (the last symbols of a pattern in one lookup; a search that fails inside them is redone step by step) and without it."""
    from gecoz_b200 import synth
    text, ref = block_1m
    og = O.GSSA(ref["gcz_body"], len(text), ref["gcx_body"])
    if not table:
        monkeypatch.setenv("GCZ_NO_KMER_TABLE", "1")
    g = G.GSSA.open(0, ref["gcz_body"], len(text), ref["gcx_body"])
    monkeypatch.delenv("GCZ_NO_KMER_TABLE", raising=False)
    data, off = synth.patterns(text, 20_000, 1, 60, seed=2)
    # edge cases appended: poly-N (huge interval), absent symbol, separator, byte >= 0x80
    extra = [b"N" * 30, b"NNNNA", b"Z", b"AC\0", b"\0", bytes([200, 65]), b"A"]
    data = np.concatenate([data, np.frombuffer(b"".join(extra), np.uint8)])
    off = np.concatenate([off, off[-1] + np.cumsum([len(x) for x in extra])]).astype(np.int64)
    sp, ep = g.count_batch(packed=(data, off))
    esp, eep, calls = og.search_batch(data, off)
    assert np.array_equal(sp, esp) and np.array_equal(ep, eep)
    assert calls > 0 and int((ep >= sp).sum()) > 9_000
    g.close()


def test_locate_rows(G, O, block_1m):
    text, ref = block_1m
    g = G.GSSA.open(0, ref["gcz_body"], len(text), ref["gcx_body"])
    rows = np.random.default_rng(3).integers(0, len(text), 50_000)
    assert np.array_equal(g.locate_rows(rows), ref["sa"][rows].astype(np.int64))
    g.close()


def _as_lists(res):
    return None if res is None else [None if x is None else np.asarray(x).tolist() for x in res]


def test_find_batch_single_string(G, O, block_1m):
    from gecoz_b200 import synth
    text, ref = block_1m
    og = O.GSSA(ref["gcz_body"], len(text), ref["gcx_body"])
    g = G.GSSA.open(0, ref["gcz_body"], len(text), ref["gcx_body"])
    data, off = synth.patterns(text, 300, 4, 14, seed=8)
    pats = [data[off[i]:off[i + 1]].tobytes() for i in range(300)] + [b"NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN", b"ZZ"]
    got = g.find_batch(pats)
    for p, r in zip(pats, got):
        assert _as_lists(r) == _as_lists(og.find(p)), p
    assert g.count(pats[0]) == [len(got[0][0])] if got[0] is not None else g.count(pats[0]) is None
    g.close()


def test_find_batch_merged_block_with_reference_quirk(G, O):
    """Merged block whose later strings sort BEFORE the first one: LF across a separator
    lands one row low in the reference; the GPU path must reproduce exactly what the oracle's literal
    LF-walk reports, whatever that is."""
    from gecoz_b200 import synth
    seqs = [synth.iid_acgtn(5000, 31), synth.iid_acgtn(3000, 32), synth.iid_acgtn(2987, 33), synth.iid_acgtn(19, 34)]
    seqs[0][:4] = np.frombuffer(b"TTTT", np.uint8)            # s1 is the largest: every other start sorts before it
    text = synth.block_of(seqs)
    ref = O.build_block(text, 32)
    og = O.GSSA(ref["gcz_body"], len(text), ref["gcx_body"])
    g = G.GSSA.open(0, ref["gcz_body"], len(text), ref["gcx_body"])
    assert g.n_strings == og.n_strings == 4
    assert g.e.tolist() == og.string_ends().tolist()
    rng = np.random.default_rng(4)
    pats = []
    for s in seqs:
        for _ in range(60):
            ln = int(rng.integers(2, 9))
            a = int(rng.integers(0, max(1, len(s) - ln)))
            pats.append(s[a:a + ln].tobytes())
        pats.append(s[:5].tobytes())                          # occurrences at the very start of a string
        pats.append(s[:2].tobytes())
    got = g.find_batch(pats)
    for p, r in zip(pats, got):
        assert _as_lists(r) == _as_lists(og.find(p)), p
    g.close()


# ---- files ---------------------------------------------------------------------------------------------------------
def test_files_match_oracle_and_roundtrip(G, O, tmp_path):
    from gecoz_b200 import synth
    recs = [(f"seq{i} some description", synth.iid_acgtn(int(ln), 40 + i)) for i, ln in
            enumerate([90_000, 61_000, 30_500, 30_000, 9_000, 500, 20, 20])]
    fa = tmp_path / "x.fa"
    with open(fa, "wb") as f:
        for h, s in recs:
            f.write(b">" + h.encode() + b"\n")
            for i in range(0, len(s), 60):
                f.write(s[i:i + 60].tobytes() + b"\r\n")
    info = G.index(fa, tmp_path / "x.gcz")
    gcz, gcx, blocks = O.write_files([(h, s.tobytes()) for h, s in recs])
    assert (tmp_path / "x.gcz").read_bytes() == gcz
    assert (tmp_path / "x.gcx").read_bytes() == gcx
    assert info["blocks"] == [[recs[i][0] for i in b] for b in blocks]
    assert len(blocks) < len(recs)                           # some sequences were merged
    with G.GecozFileReader(tmp_path / "x.gcz") as reader:
        assert G.GecozFileReader.checkFormat(tmp_path / "x.gcz")
        hdr = reader.findBlockHeader("seq3 some description")
        assert hdr is not None and reader.findBlockHeader("seq3") is None     # full header line, exact match
        ssa = reader.read(hdr)
        nstr = hdr.findHeader("seq3 some description")
        pat = recs[3][1][1000:1012].tobytes()
        res = ssa.find(pat)
        assert res is not None and 1000 in res[nstr].tolist()
        assert ssa.getLength(nstr) == 30_000
        ssa.close()


def test_missing_gcx_fails_fast(G, O, tmp_path):
    from gecoz_b200 import synth
    G.index_records([("a", synth.iid_acgtn(5000, 1))], tmp_path / "a.gcz")
    (tmp_path / "a.gcx").unlink()
    with G.GecozFileReader(tmp_path / "a.gcz") as reader:
        with pytest.raises(G.GczError):
            reader.read(reader.getBlockHeaders()[0])


def test_corrupt_index_is_rejected(G, O, tmp_path):
    from gecoz_b200 import synth
    G.index_records([("a", synth.iid_acgtn(5000, 1))], tmp_path / "a.gcz")
    raw = bytearray((tmp_path / "a.gcx").read_bytes())
    raw[20] ^= 0xFF                                           # header hash
    (tmp_path / "a.gcx").write_bytes(raw)
    with G.GecozFileReader(tmp_path / "a.gcz") as reader:
        with pytest.raises(G.GczFormatError):
            reader.read(reader.getBlockHeaders()[0])


def test_argument_errors(G):
    text = np.frombuffer(b"ACGT\0", np.uint8).copy()
    shape = G.shape_from_counts(np.bincount(text, minlength=256))
    gcz = np.zeros(shape.size, np.uint8)
    gcx = np.zeros(G.index_size(len(text), 5), np.uint8)
    with pytest.raises(G.GczError):                          # wrong body size
        G._native.check(G.lib().gcz_build_block(0, _p(text), len(text), 32, C.byref(shape), _p(gcz), shape.size - 1,
                                                _p(gcx), len(gcx), None, None))
    with pytest.raises(G.GczError):                          # sampling rate not a power of two
        G._native.check(G.lib().gcz_build_block(0, _p(text), len(text), 24, C.byref(shape), _p(gcz), shape.size,
                                                _p(gcx), len(gcx), None, None))
    other = np.frombuffer(b"AAAA\0", np.uint8).copy()
    with pytest.raises(G.GczError):                          # shape of another text
        G._native.check(G.lib().gcz_build_block(0, _p(other), len(other), 32, C.byref(shape), _p(gcz), shape.size,
                                                _p(gcx), len(gcx), None, None))


# ---- multi-GPU partitioning (gecoz_b200/sharding.py) with the CUDA engine ---------------------------------------------
def test_sharded_index_one_rank(G, O, tmp_path):
    from gecoz_b200 import sharding, synth
    recs = [(f"s{i}", synth.iid_acgtn(int(ln), 90 + i)) for i, ln in enumerate([50_000, 33_000, 17_000, 16_000, 900, 900, 12])]
    info = sharding.sharded_index_records(recs, tmp_path / "y.gcz", engine=sharding.GpuEngine(0))
    gcz, gcx, blocks = O.write_files([(h, s.tobytes()) for h, s in recs])
    assert (tmp_path / "y.gcz").read_bytes() == gcz and (tmp_path / "y.gcx").read_bytes() == gcx
    assert info["mine"] == list(range(len(blocks)))


def _two_gpus():
    import torch
    return torch.cuda.device_count() >= 2


@pytest.mark.parametrize("mode", ["build_gpu", "query_gpu"])
def test_sharded_two_gpus_nccl(G, O, tmp_path, mode):
    """Two ranks, two GPUs, NCCL: one file pair written by both / one batch answered by both."""
    if not _two_gpus():
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import subprocess
    import sys
    here = Path(__file__).resolve().parent
    sys.path.insert(0, str(here))
    import dist_worker as W
    init = tmp_path / "rendezvous"
    procs = [subprocess.Popen([sys.executable, str(here / "dist_worker.py"), str(r), "2", str(init), str(tmp_path), mode],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    for p in procs:
        out, _ = p.communicate(timeout=600)
        assert p.returncode == 0, out
    if mode == "build_gpu":
        recs = W.records()
        gcz, gcx, _ = O.write_files([(h, s.tobytes()) for h, s in recs])
        assert (tmp_path / "x.gcz").read_bytes() == gcz and (tmp_path / "x.gcx").read_bytes() == gcx
    else:
        got = np.load(tmp_path / "query.npz")
        texts, data, off = W.query_inputs()
        for b, t in enumerate(texts):
            og = W.OracleGSSA(t)
            sp, ep = og.count_batch((data, off))
            assert np.array_equal(got["sp"][b], sp) and np.array_equal(got["ep"][b], ep)
            per, pos, poff = og.find_batch_raw((data, off))
            assert np.array_equal(got[f"per{b}"], per) and np.array_equal(got[f"pos{b}"], pos) and np.array_equal(got[f"off{b}"], poff)


# ---- callers: GecoMatch (-c / -s) and SimpleGFFGenerator (-s patterns.fa) ------------------------------------------------
def _small_genome(tmp_path, G):
    from gecoz_b200 import synth
    recs = [(f"chr{i} test", synth.iid_acgtn(int(ln), 60 + i, p_n=0.0)) for i, ln in enumerate([40_000, 25_000, 12_000, 11_000, 700, 650])]
    info = G.index_records(recs, tmp_path / "g.gcz")
    return recs, info


def _oracle_blocks(O, recs, info):
    from gecoz_b200 import synth
    by = {h: s for h, s in recs}
    out = []
    for headers in info["blocks"]:
        text = synth.block_of([by[h] for h in headers])
        r = O.build_block(text, 32)
        out.append((headers, O.GSSA(r["gcz_body"], len(text), r["gcx_body"])))
    return out


def test_geco_match_count_and_search(G, O, tmp_path):
    from gecoz_b200 import geco_match
    recs, info = _small_genome(tmp_path, G)
    blocks = _oracle_blocks(O, recs, info)
    assert len(blocks) < len(recs)                                       # merged blocks: per-string split matters
    pats = [recs[0][1][100:108].tobytes(), recs[3][1][5:11].tobytes(), recs[5][1][:4].tobytes(), b"ACGTACGTACGTACGTACGTAC", b"A"]
    for pat in pats:
        for want_pos in (False, True):
            exp = []
            for headers, og in blocks:
                res = og.find(pat)
                if res is None:
                    continue
                for h, r in zip(headers, res):
                    if r is not None and len(r):
                        exp.append(f">{h} found : {len(r)}")
                        if want_pos:
                            exp += [str(int(p)) for p in r]
            got = (geco_match.match if want_pos else geco_match.count)(tmp_path / "g.gcz", None, pat)
            assert got == exp, pat
    # -s / -c with a header: only that string of that block
    hdr = recs[3][0]
    headers, og = next(b for b in blocks if hdr in b[0])
    res = og.find(pats[1])
    r = res[headers.index(hdr)] if res is not None else None
    exp = ([f">{hdr} found : {len(r)}"] + [str(int(p)) for p in r]) if r is not None and len(r) else []
    assert geco_match.match(tmp_path / "g.gcz", hdr, pats[1]) == exp
    with pytest.raises(KeyError):
        geco_match.count(tmp_path / "g.gcz", "chrZ", b"ACGT")


def test_simple_gff_generator(G, O, tmp_path):
    from gecoz_b200 import geco_match
    recs, info = _small_genome(tmp_path, G)
    blocks = _oracle_blocks(O, recs, info)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    s0, s2 = recs[0][1], recs[2][1]
    pats = [("p1|first note|second", s0[200:230].tobytes()), ("rna", s0[300:320].tobytes().replace(b"T", b"U")),
            ("revcomp|", s2[50:75].tobytes().translate(comp)[::-1]), ("absent", b"ACGT" * 9), ("|", s2[10:22].tobytes()),
            ("short", b"ACG")]
    fa = tmp_path / "p.fa"
    with open(fa, "wb") as f:
        for i, (h, s) in enumerate(pats):
            if i == 1:                                                    # a FASTQ record in the middle
                f.write(b"@" + h.encode() + b"\r\n" + s[:10] + b"\r\n" + s[10:] + b"\n+\n" + b"I" * len(s) + b"\n")
            else:
                f.write(b">" + h.encode() + b"\n" + s + b"\n")
        f.write(b">empty\n>last\nACGTAC\n")
    exp = []
    for h, s in pats + [("last", b"ACGTAC")]:
        fwd = s.replace(b"U", b"T")
        for strand, seq in (("+", fwd), ("-", fwd.translate(comp)[::-1])):
            for headers, og in blocks:
                res = og.find(seq)
                if res is None:
                    continue
                for name, r in zip(headers, res):
                    if r is None:
                        continue
                    parts = h.split("|")
                    while parts and parts[-1] == "":
                        parts.pop()
                    attrs = ("ID=" + parts[0] if parts else "") + "".join(";Note=" + x for x in parts[1:])
                    exp += [f"{name}\tgecotools\tdna\t{int(p) + 1}\t{int(p) + len(seq)}\t1.000\t{strand}\t.\t{attrs}" for p in r]
    got = geco_match.search(tmp_path / "g.gcz", fa)
    assert got == exp
    assert any("\t-\t" in x for x in got) and any("\t+\t" in x for x in got) and len(got) > 10


# ---- extract (GSSA.extract, GSSAIndex.find, IndexWaveletTree.find, select) and GecoRead -----------------------------------
def test_extract_single_string(G, O):
    from gecoz_b200 import synth
    seq = synth.chromosome_shaped(300_000, 9)
    text = synth.block_of([seq])
    ref = O.build_block(text, 32)
    g = G.GSSA.open(0, ref["gcz_body"], len(text), ref["gcx_body"])
    og = O.GSSA(ref["gcz_body"], len(text), ref["gcx_body"])
    rng = np.random.default_rng(3)
    cases = [(0, len(seq)), (0, len(seq) + 100), (0, 1), (31, 2), (32, 32), (33, 31), (len(seq) - 1, 5), (len(seq) - 40, 40),
             (12345, 70_000)] + [(int(a), int(b)) for a, b in zip(rng.integers(0, len(seq), 20), rng.integers(1, 5000, 20))]
    for start, cap in cases:
        got = g.extract(0, start, cap)
        exp = og.extract(0, start, cap)
        assert np.array_equal(got, exp), (start, cap)
        assert np.array_equal(got, seq[start:start + cap])                 # single-string blocks extract exactly
    with pytest.raises(IndexError):
        g.extract(1, 0, 10)
    g.close()


@pytest.mark.parametrize("rate", [32, 4])
def test_extract_merged_block_like_the_reference(G, O, rate):
    """Merged block: calls whose walk starts beyond the string end cross a separator the LF mapping gets wrong
     — the bytes must be the reference's (the oracle's literal walk), right or wrong."""
    from gecoz_b200 import synth
    seqs = [synth.iid_acgtn(5000, 31), synth.iid_acgtn(3000, 32), synth.iid_acgtn(2987, 33), synth.iid_acgtn(19, 34),
            synth.iid_acgtn(64, 35), synth.iid_acgtn(1, 36)]
    seqs[0][:4] = np.frombuffer(b"TTTT", np.uint8)
    text = synth.block_of(seqs)
    ref = O.build_block(text, rate)
    g = G.GSSA.open(0, ref["gcz_body"], len(text), ref["gcx_body"])
    og = O.GSSA(ref["gcz_body"], len(text), ref["gcx_body"])
    differs = 0
    for nstr, s in enumerate(seqs):
        for start, cap in ((0, len(s) + 10), (0, len(s)), (0, 100), (17, 64), (max(len(s) - 5, 0), 100), (len(s) // 2, 1), (0, 31)):
            if start >= len(s):
                continue
            try:
                exp = og.extract(nstr, start, cap)
            except ValueError:                                              # Java: IllegalArgumentException
                with pytest.raises(G.GczError):
                    g.extract(nstr, start, cap)
                differs += 1
                continue
            got = g.extract(nstr, start, cap)
            assert np.array_equal(got, exp), (nstr, start, cap)
            differs += not np.array_equal(got, s[start:start + cap])
    assert differs > 0 or rate != 32                                        # the quirk is exercised at rate 32
    g.close()


def test_geco_read_fasta_and_sequence(G, O, tmp_path):
    from gecoz_b200 import geco_read
    recs, info = _small_genome(tmp_path, G)
    blocks = _oracle_blocks(O, recs, info)
    n = geco_read.fasta(tmp_path / "g.gcz", tmp_path / "out.fa")
    assert n == len(recs)
    exp = b""
    for headers, og in blocks:
        for nstr, h in enumerate(headers):
            length = int(og.string_ends()[nstr] - (og.string_ends()[nstr - 1] + 1 if nstr else 0))
            seq = og.extract(nstr, 0, 4 * 1024 * 1024)[:length]
            body = bytearray()
            for i in range(0, length, 50):
                body += seq[i:i + 50].tobytes() + b"\n"
            if length % 50 == 0:
                body += b"\n"
            exp += b">" + h.encode() + b"\n" + bytes(body)
    assert (tmp_path / "out.fa").read_bytes() == exp
    hdr, seq = recs[1]
    assert geco_read.sequence(tmp_path / "g.gcz", hdr, 100, 1100, tmp_path / "sub.bin") == 1000
    blk = next(b for b in blocks if hdr in b[0])
    assert (tmp_path / "sub.bin").read_bytes() == blk[1].extract(blk[0].index(hdr), 100, 1000).tobytes()


# ---- the native host layer (include/gcz_file.h) with the CUDA engine -------------------------------------------------------
def test_native_writer_and_reader_on_the_gpu(G, O, tmp_path):
    from gecoz_b200 import native_file as NF, synth
    recs = [(f"n{i} d", synth.iid_acgtn(int(ln), 80 + i)) for i, ln in enumerate([30_000, 21_000, 9_500, 9_000, 300, 300, 11])]
    fa = tmp_path / "n.fa"
    with open(fa, "wb") as f:
        for h, s in recs:
            f.write(b">" + h.encode() + b"\n")
            for i in range(0, len(s), 70):
                f.write(s[i:i + 70].tobytes() + b"\n")
    with NF.Fasta(fa) as fasta:
        rep = NF.index(fasta, tmp_path / "n.gcz")                            # engine = the library's CUDA entry points
    gcz, gcx, blocks = O.write_files([(h, s.tobytes()) for h, s in recs])
    assert (tmp_path / "n.gcz").read_bytes() == gcz and (tmp_path / "n.gcx").read_bytes() == gcx
    assert rep["blocks"] == len(blocks)
    with NF.Reader(tmp_path / "n.gcz") as r:
        b, s = r.find(recs[2][0])
        g = r.open_block(b)
        pat = recs[2][1][100:120].tobytes()
        res = g.find(pat)
        assert res is not None and res[s] is not None and 100 in res[s].tolist()
        assert np.array_equal(g.extract(s, 50, 500), recs[2][1][50:550])
        g.close()


def test_native_callers_on_the_gpu(G, O, tmp_path):
    """gcz_match / gcz_gff_search / gcz_extract_fasta with the library's own CUDA entry points as the engine."""
    from gecoz_b200 import geco_match, geco_read, native_file as NF
    recs, info = _small_genome(tmp_path, G)
    pat = recs[0][1][100:112].tobytes()
    fa = tmp_path / "p.fa"
    fa.write_bytes(b">p1|note\n" + recs[2][1][50:80].tobytes() + b"\n>p2\n" + recs[0][1][10:40].tobytes() + b"\n")
    with NF.Reader(tmp_path / "g.gcz") as r:
        assert r.match(None, pat, True) == geco_match.match(tmp_path / "g.gcz", None, pat)
        assert r.match(recs[0][0], pat, False) == geco_match.count(tmp_path / "g.gcz", recs[0][0], pat)
        assert r.gff_search(fa.read_bytes()) == geco_match.search(tmp_path / "g.gcz", fa)
        assert r.extract_fasta(tmp_path / "native.fa") == len(recs)
    geco_read.fasta(tmp_path / "g.gcz", tmp_path / "python.fa")
    assert (tmp_path / "native.fa").read_bytes() == (tmp_path / "python.fa").read_bytes()


# ---- a batch against every block of a file: gcz_count_multi / gcz_find_multi ------------------------------------------------
@pytest.fixture(scope="module")
def three_blocks(G, O):
    """Three blocks as GecoIndex would write them: one single-string block, one merged block whose later strings sort before the
    first one (the LF-across-separator quirk) and one tiny block."""
    from gecoz_b200 import synth
    blocks = [[synth.iid_acgtn(200_000, 41)],
              [synth.iid_acgtn(5000, 31), synth.iid_acgtn(3000, 32), synth.iid_acgtn(2987, 33), synth.iid_acgtn(19, 34)],
              [synth.iid_acgtn(300, 51), synth.iid_acgtn(41, 52)]]
    blocks[1][0][:4] = np.frombuffer(b"TTTT", np.uint8)
    out = []
    for seqs in blocks:
        text = synth.block_of(seqs)
        ref = O.build_block(text, 32)
        out.append((text, ref, G.GSSA.open(0, ref["gcz_body"], len(text), ref["gcx_body"]), O.GSSA(ref["gcz_body"], len(text), ref["gcx_body"])))
    yield out
    for _, _, g, og in out:
        g.close()
        og.close()


def _mixed_patterns(three_blocks, count=4000, seed=3):
    from gecoz_b200 import synth
    rng = np.random.default_rng(seed)
    pats = []
    for text, _, _, _ in three_blocks:
        for _ in range(count // 3):
            ln = int(rng.integers(1, 24))
            a = int(rng.integers(0, max(1, len(text) - ln)))
            pats.append(text[a:a + ln].tobytes())              # may span a separator: patterns holding '\0' hit string ends exactly
    pats += [b"A", b"N", b"\0", b"T\0", b"\0A", b"ZZ", bytes([200, 65]), b"ACGT" * 30]
    return pats


def test_count_multi_equals_the_sum_over_blocks(G, three_blocks):
    import torch
    pats = _mixed_patterns(three_blocks)
    data, off = G.pack_patterns(pats)
    gssas = [g for _, _, g, _ in three_blocks]
    exp = np.zeros(len(pats), np.int64)
    for g in gssas:
        sp, ep = g.count_batch(packed=(data, off))
        exp += np.maximum(ep - sp + 1, 0)
    assert np.array_equal(G.count_totals(gssas, data, off), exp)
    # device-resident batch and result, and a pinned host batch
    d_data, d_off = torch.from_numpy(data).cuda(), torch.from_numpy(off).cuda()
    d_out = torch.full((len(pats),), -7, dtype=torch.int64, device="cuda")
    G.count_totals(gssas, d_data, d_off, d_out)
    assert np.array_equal(d_out.cpu().numpy(), exp)
    h_out = torch.empty(len(pats), dtype=torch.int64).pin_memory()
    G.count_totals(gssas, torch.from_numpy(data).pin_memory(), torch.from_numpy(off).pin_memory(), h_out)
    assert np.array_equal(h_out.numpy(), exp)
    st = G.last_query_stats()
    assert st["patterns"] == len(pats) and st["blocks"] == 3 and st["kernel_ms"] > 0


def test_count_stats_equal_the_instrumented_oracle(G, three_blocks):
    pats = _mixed_patterns(three_blocks, count=1500, seed=5)
    data, off = G.pack_patterns(pats)
    calls = 0
    for _, _, g, og in three_blocks:
        calls += og.search_batch(data, off)[2]
    st = G.count_stats([g for _, _, g, _ in three_blocks], data, off)
    assert st["reference_rank_calls"] == calls
    assert 0 < st["rank_sectors"] <= calls and st["steps"] > 0 and st["index_bytes"] > 0


def test_find_multi_equals_the_oracles_find_per_block(G, three_blocks):
    """Sparse hit records of a batch over three blocks == GSSA.find of the oracle, pattern by pattern and block by block —
    including patterns that hold a separator (their positions ARE string ends: the reference's binarySearch finds the key
    and takes its other branch) and the merged block whose LF walk crosses separators."""
    import torch
    pats = _mixed_patterns(three_blocks)
    data, off = G.pack_patterns(pats)
    gssas = [g for _, _, g, _ in three_blocks]
    for src in ((data, off), (torch.from_numpy(data).cuda(), torch.from_numpy(off).cuda())):
        block_off, pattern, string, position = G.find_multi(gssas, *src)
        assert block_off[0] == 0 and block_off[-1] == len(pattern) == len(string) == len(position)
        for b, (_, _, g, og) in enumerate(three_blocks):
            lo, hi = int(block_off[b]), int(block_off[b + 1])
            pb, sb, xb = pattern[lo:hi], string[lo:hi], position[lo:hi]
            assert (np.diff(pb) >= 0).all()
            starts = np.searchsorted(pb, np.arange(len(pats) + 1))
            for i, p in enumerate(pats):
                if len(p) == 0 or max(p) >= 128:
                    exp = None
                else:
                    exp = og.find(p)
                got = [[] for _ in range(g.n_strings)]
                for j in range(starts[i], starts[i + 1]):
                    got[int(sb[j])].append(int(xb[j]))
                want = [[] if (exp is None or x is None) else x.tolist() for x in (exp or [None] * g.n_strings)]
                assert got == want, (b, p)


def test_find_batch_dense_form_still_matches(G, three_blocks):
    pats = _mixed_patterns(three_blocks, count=900, seed=9)
    for _, _, g, og in three_blocks:
        got = g.find_batch(pats)
        for p, r in zip(pats, got):
            if max(p) >= 128:
                continue
            assert _as_lists(r) == _as_lists(og.find(p)), p


def test_build_refuses_a_shape_from_another_histogram(G):
    """Same alphabet, other counts: the file slices the caller reserved would not fit what the text needs (the reference
    trusts its caller here and corrupts the file)."""
    from gecoz_b200 import synth
    a = synth.cfg2_text(100_000, seed=41)
    b = a.copy()
    b[1000:3000] = ord("A")                                   # same symbols present, other histogram
    shape_b = G.shape_from_counts(np.bincount(b, minlength=256).astype(np.int64))
    shape_a = G.shape_from_counts(np.bincount(a, minlength=256).astype(np.int64))
    if shape_a.size == shape_b.size:
        pytest.skip("the two histograms happen to give the same sizes")
    gcz = np.zeros(shape_b.size, np.uint8)
    gcx = np.zeros(G.index_size(len(a), 5), np.uint8)
    with pytest.raises(G.GczError, match="shape does not match"):
        G.build_block(0, a, len(a), 32, shape_b, gcz, gcx)


def test_find_in_small_chunks_and_with_a_grown_workspace(G, three_blocks):
    """The two paths of find that ordinary batches do not reach: occurrences cut into several locate / sort / split launches
    (here 1000 per launch instead of 2^26; a pattern with more than that gets a launch of its own), and a batch whose occurrences
    outgrow the workspace reserved up front (the block is redone after the arena has grown)."""
    pats = _mixed_patterns(three_blocks, count=1500, seed=12)
    data, off = G.pack_patterns(pats)
    gssas = [g for _, _, g, _ in three_blocks]
    whole = G.find_multi(gssas, data, off)
    G._native.check(G.lib().gcz_dbg_set_find_chunk(1000))
    try:
        chunked = G.find_multi(gssas, data, off)
    finally:
        G._native.check(G.lib().gcz_dbg_set_find_chunk(0))
    for a, b in zip(whole, chunked):
        assert np.array_equal(a, b)
    # 150 x "A": ~50 000 occurrences each in the first block, 7.5 M in all = 270 MB of occurrence workspace
    text, _, g, og = three_blocks[0]
    many = [b"A"] * 150
    data, off = G.pack_patterns(many)
    block_off, pattern, string, position = G.find_multi([g], data, off)
    exp = og.find(b"A")[0]
    assert len(position) == 150 * len(exp) and block_off.tolist() == [0, len(position)]
    assert np.array_equal(pattern, np.repeat(np.arange(150), len(exp)))
    assert np.array_equal(position.reshape(150, -1), np.tile(exp, (150, 1))) and not string.any()
