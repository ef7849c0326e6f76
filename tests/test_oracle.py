"""Pins the CPU oracle (oracle/) before anything is compared against it.

Sources of truth, in decreasing strength:
  1. the reference's own known-answer tests / documentation table (tests/golden/reference_kats.json),
  2. the reference's property tests restated (DeflateTablesTest.java:57-198),
  3. an independent transliteration's whole-block bytes (tests/golden/provisional_blocks.json),
  4. invariants from SURVEY.md App. B.9 (sizes, re-open, find == naive search).
"""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import gcz_oracle as O

GOLD = Path(__file__).parent / "golden"
KAT = json.loads((GOLD / "reference_kats.json").read_text())
PROV = json.loads((GOLD / "provisional_blocks.json").read_text())

# DeflateTablesTest.java:42-55
TEXT = ("My uncle, what a worthy man, Falling ill like that, and dying; It summons up respect, one can "
        "Admire it, as if he were trying. Let us all follow his example! But, God, what tedium to sample "
        "That sitting by the bed all day, All night, barely a foot away! And the hypocrisy, demeaning, "
        "Of cosseting one who's half alive; Puffing the pillows, you contrive To bring his medicine unsmiling, "
        "Thinking with a mournful sigh, 'Why the devil can't you die?'").encode()


# ---- 1. reference known answers ---------------------------------------------------------------
def test_bitbuffer_write_kat():
    k = KAT["bitbuffer_write"]
    data, pos = O.bitbuffer_write([int(k["value"], 2)], [k["nbits"]], k["cap"])
    assert list(data[:3]) == k["bytes"] and pos == 3


def test_bitbuffer_flush_into_short_buffer_kat():
    k = KAT["bitbuffer_flush_short"]
    data, pos = O.bitbuffer_write([int(k["value"], 2)], [k["nbits"]], k["cap"])
    assert list(data) == k["bytes"] and pos == 3


def test_bitbuffer_read_kat():
    k = KAT["bitbuffer_read"]
    out = O.bitbuffer_write_read([int(k["value"], 2)], [k["nbits"]], 128, [k["read_nbits"]])
    assert int(out[0]) & 15 == k["expect_low4"]


def test_iwt_pdf_table3():
    k = KAT["iwt_table3"]
    vals = k["values"]
    m = len(vals)
    buf = O.iwt_write(vals)
    assert buf.tobytes().hex() == k["serialized_hex"]
    nb = O.ranked_bytes(m)
    for lvl, row in enumerate(k["levels_high_to_low"]):
        node = buf[lvl * nb:(lvl + 1) * nb]
        assert "".join(str(O.ranked_get(node, m, i)) for i in range(m)) == row
    assert [O.iwt_get(buf, m, i) for i in range(m)] == vals
    assert [O.iwt_find(buf, m, v) for v in vals] == list(range(m))


# ---- 2. reference property tests ----------------------------------------------------------------
def _counts(data: bytes):
    return np.bincount(np.frombuffer(data, np.uint8), minlength=256).astype(np.int64)


def _stress_data() -> bytes:
    return b"".join(bytes([i]) * (i * i + 1) for i in range(256))     # DeflateTablesTest.java:140-146


@pytest.mark.parametrize("data", [TEXT, _stress_data()], ids=["text", "stress"])
def test_code_gens(data):
    c = _counts(data)
    bl, tb = O.deflate_encode_table(c)
    assert all((c[i] == 0) == (bl[i] == 0) for i in range(256))
    assert bl.max() <= 15
    for i in range(256):
        if c[i] > 0:
            assert O.lookup_get_symbol(bl, int(tb[i])) == i
    # prefix-free (Kraft equality for a full tree)
    assert sum(2.0 ** -int(l) for l in bl if l > 0) == 1.0


@pytest.mark.parametrize("data,cap", [(TEXT, 252), (_stress_data(), None)], ids=["text", "stress"])
def test_stream_roundtrip(data, cap):
    assert O.deflate_stream_roundtrip(data, cap or len(data)) == 0


def test_long_codes_name_nodes_injectively():
    """Codes > 9 bits go through the lookup table's extension area; the HSWT needs distinct names."""
    c = _counts(_stress_data())
    bl, tb = O.deflate_encode_table(c)
    assert bl.max() > 9
    names = {}
    for s in range(256):
        for j in range(int(bl[s])):
            prefix = int(tb[s]) & ((1 << j) - 1)
            names[(prefix, j)] = O.lookup_get_symbol(bl, prefix | (1 << j))
    assert len(set(names.values())) == len(names) == 255


# ---- 3. independent transliteration ----------------------------------------------------------------
@pytest.mark.parametrize("name", ["gattaca", "two_strings"])
def test_provisional_blocks(name):
    p = PROV[name]
    text = b"".join(s.encode() + b"\0" for s in p["sequences"])
    r = O.build_block(text, p["sampling_rate"], want_sa=True, want_bwt=True)
    assert r["sa"].tolist() == p["sa"]
    assert r["bwt"].tobytes().hex() == p["bwt_hex"]
    hl = 26 + sum(len(h) + 1 for h in p["headers"])
    gcz = O.ref_header(p["headers"], hl + len(r["gcz_body"]), len(text)) + r["gcz_body"].tobytes()
    gcx = O.ssa_header(p["headers"], len(r["gcx_body"])) + r["gcx_body"].tobytes()
    assert gcz.hex() == p["gcz_hex"]
    assert gcx.hex() == p["gcx_hex"]
    g = O.GSSA(r["gcz_body"], len(text), r["gcx_body"])
    for pat, exp in p["find"].items():
        got = g.find(pat.encode())
        assert [None if x is None else x.tolist() for x in got] == exp


# ---- 4. invariants ----------------------------------------------------------------------------------
def _rand_text(rng, n, alphabet=b"ACGT", p_n=0.01, nseq=1):
    parts = []
    for _ in range(nseq):
        a = np.frombuffer(alphabet, np.uint8)[rng.integers(0, len(alphabet), n)]
        a = a.copy()
        a[rng.random(n) < p_n] = ord("N")
        parts.append(a.tobytes() + b"\0")
    return b"".join(parts)


@pytest.mark.parametrize("seed", range(6))
def test_sais_equals_naive(seed):
    rng = np.random.default_rng(seed)
    kind = seed % 3
    if kind == 0:
        t = _rand_text(rng, 3000, nseq=3)
    elif kind == 1:
        t = bytes(rng.integers(0, 256, 5000, dtype=np.uint8))          # full byte alphabet incl. >= 0x80
    else:
        t = (b"ACACACAC" * 200 + b"N" * 700 + b"\0" + b"A" * 500 + b"\0\0" + b"CA" * 300 + b"\0")
    assert np.array_equal(O.suffix_array(t), O.suffix_array(t, naive=True))


@pytest.mark.parametrize("n", [1, 7, 63, 64, 65, 511, 512, 513, 4096, 65535, 65536, 65537, 70000, 140000])
def test_ranked_vector_layout(n):
    rng = np.random.default_rng(n)
    bits = (rng.random(n) < 0.37).astype(np.uint8)
    buf = O.ranked_write(bits)
    assert len(buf) == O.ranked_bytes(n) == ((n - 1) >> 16) * 6 + ((n - 1) >> 9) * 2 + ((n + 7) >> 3)
    cs = np.cumsum(bits)
    for i in sorted({0, n - 1, n // 2, min(n - 1, 511), min(n - 1, 512), min(n - 1, 65535), min(n - 1, 65536)} |
                    set(rng.integers(0, n, 40).tolist())):
        assert O.ranked_get(buf, n, i) == bits[i]
        assert O.ranked_count(buf, n, i) == cs[i]
    ones = np.flatnonzero(bits)
    zeros = np.flatnonzero(bits == 0)
    for k in rng.integers(0, len(ones), 10) if len(ones) else []:
        assert O.ranked_find_one(buf, n, int(k) + 1) == ones[k]
    for k in rng.integers(0, len(zeros), 10) if len(zeros) else []:
        assert O.ranked_find_zero(buf, n, int(k) + 1) == zeros[k]


def _naive_find(seqs, pat: bytes):
    res = []
    for s in seqs:
        pos, i = [], s.find(pat)
        while i >= 0:
            pos.append(i)
            i = s.find(pat, i + 1)
        res.append(pos or None)
    return res if any(r for r in res) else None


@pytest.mark.parametrize("seed,n,rate", [(0, 2000, 32), (1, 70000, 32), (2, 5000, 4), (3, 1500, 2)])
def test_block_invariants_single_string(seed, n, rate):
    rng = np.random.default_rng(100 + seed)
    text = _rand_text(rng, n)
    r = O.build_block(text, rate, want_sa=True, want_bwt=True)
    s = r["shape"]
    # reserved sizes are exact: a canary build into the same sizes succeeded, and the last node byte is data
    assert len(r["gcz_body"]) == s.size and len(r["gcx_body"]) == O.index_size(len(text), rate.bit_length() - 1)
    g = O.GSSA(r["gcz_body"], len(text), r["gcx_body"])
    # re-open recovers every node length (mapNodes via count(len-1)) in file order
    names, lens, offs = g.nodes()
    assert [int(s.node_bits[nm]) for nm in names] == lens.tolist()
    assert offs[0] == s.table_bytes
    assert offs[-1] + O.ranked_bytes(int(lens[-1])) == s.size
    if O.index_size(len(text), rate.bit_length() - 1) > O.index_size(len(text), rate.bit_length()):
        assert g.sampling_factor == rate.bit_length() - 1
    # C array and occ
    t = np.frombuffer(text, np.uint8)
    cnt = np.bincount(t, minlength=256)
    assert g.c_array().tolist() == (np.cumsum(cnt) - cnt).tolist()
    bwt = r["bwt"]
    for sym in (0, ord("A"), ord("N"), ord("T"), ord("Z")):
        for pos in (0, 1, len(text) // 3, len(text) - 1):
            exp = int((bwt[:pos + 1] == sym).sum()) - 1
            assert g.occ(sym, pos) == exp
    # locate == SA, LF walk
    sa = r["sa"]
    for row in rng.integers(0, len(text), 50):
        assert g.locate(int(row)) == sa[row]
    assert g.string_ends().tolist() == [len(text) - 1]
    # find == naive
    seq = text[:-1]
    pats = [seq[i:i + l] for i, l in zip(rng.integers(0, n - 20, 20), rng.integers(1, 20, 20))]
    pats += [bytes(rng.choice(list(b"ACGT"), 6)) for _ in range(10)] + [b"N", b"A", b"ZZ"]
    for p in pats:
        got = g.find(p)
        exp = _naive_find([seq], p)
        assert (None if got is None else [None if x is None else x.tolist() for x in got]) == exp
        sp, ep, calls = g.search(p)
        assert max(0, ep - sp + 1) == (len(exp[0]) if exp else 0)


def test_merged_block_find_benign():
    """Strings sorted so that every later string sorts AFTER the first one: LF across separators is exact."""
    seqs = [b"AACGTACGTTAGC" * 7, b"CCGTA" * 9, b"GGT" * 11]     # s1 < s2 < s3 lexicographically
    text = b"".join(s + b"\0" for s in seqs)
    r = O.build_block(text, 4)
    g = O.GSSA(r["gcz_body"], len(text), r["gcx_body"])
    assert g.n_strings == 3
    assert g.string_ends().tolist() == [len(seqs[0]), len(seqs[0]) + 1 + len(seqs[1]), len(text) - 1]
    for p in (b"CGT", b"GT", b"TA", b"GGTGG", b"AAC", b"TTTT"):
        got = g.find(p)
        exp = _naive_find(seqs, p)
        assert (None if got is None else [None if x is None else x.tolist() for x in got]) == exp


def test_iwt_roundtrip_random():
    rng = np.random.default_rng(5)
    for m in (1, 2, 3, 17, 512, 513, 5000):
        vals = rng.permutation(m).astype(np.int32)
        buf = O.iwt_write(vals)
        assert len(buf) == O.ranked_bytes(m) * m.bit_length()
        idx = rng.integers(0, m, min(m, 64))
        assert [O.iwt_get(buf, m, int(i)) for i in idx] == vals[idx].tolist()


def test_general_alphabet_block():
    rng = np.random.default_rng(9)
    # lower case, IUPAC, skew: more than 9-bit codes are not reached here but > 6 symbols are
    alpha = b"ACGTNacgtnRYKMSW"
    p = np.array([30, 30, 30, 30, 3, 10, 10, 10, 10, 1, .5, .5, .2, .2, .1, .1])
    t = np.frombuffer(alpha, np.uint8)[rng.choice(len(alpha), 20000, p=p / p.sum())].tobytes() + b"\0"
    r = O.build_block(t, 32, want_sa=True)
    g = O.GSSA(r["gcz_body"], len(t), r["gcx_body"])
    for p_ in (b"ACg", b"nn", b"RY", t[100:130], t[5000:5008]):
        got = g.find(p_)
        exp = _naive_find([t[:-1]], p_)
        assert (None if got is None else [None if x is None else x.tolist() for x in got]) == exp


def test_header_and_hash():
    assert O.header_hash(["s1"]) == int.from_bytes(bytes.fromhex("c3a8ffffffff030f"), "little", signed=True)
    h = O.ref_header(["chr1", "chrM"], 1234, 77)
    assert h == b"GecozBWT\x01" + (1234).to_bytes(8, "little") + (77).to_bytes(8, "little") + b"chr1\0chrM\0\0"
    assert O.ssa_header(["a"], 9)[:17] == b"GecozSSA\x01" + (9).to_bytes(8, "little")


HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
        133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616,
        64444167, 46709983, 50818468, 156040895, 57227415, 16569]
HG38_NAMES = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY", "chrM"]


def test_merge_blocks_hg38_layout():
    """18 blocks from simulating tools/GecoIndex.java:72-98."""
    blocks = O.merge_blocks(HG38, HG38_NAMES)
    got = [[HG38_NAMES[i] for i in b] for b in blocks]
    exp = [["chr1"], ["chr2"], ["chr3"], ["chr4"], ["chr5"], ["chr6"], ["chr7"], ["chrX"], ["chr8"], ["chr9"],
           ["chr11"], ["chr10"], ["chr12"], ["chr13", "chr14"], ["chr15", "chr22", "chr21", "chrM"],
           ["chr16", "chr17"], ["chr18", "chr20"], ["chr19", "chrY"]]
    assert got == exp
    sizes = [sum(HG38[i] + 1 for i in b) for b in blocks]
    assert sizes[:3] == [248956423, 242193530, 198295560] and sum(sizes) == 3088286426


def test_write_files_two_blocks():
    gcz, gcx, blocks = O.write_files([("a", b"ACGTN"), ("b", b"ACG")], sampling_rate=2)
    assert blocks == [[0], [1]]
    assert gcz.count(b"GecozBWT") == 2 and gcx.count(b"GecozSSA") == 2
    size0 = int.from_bytes(gcz[9:17], "little")
    assert gcz[size0:size0 + 8] == b"GecozBWT"


# ---- extract (GSSA.extract :90-126, GSSAIndex.find :184-187) ------------------------------------------------------------
def test_oracle_extract_and_index_find():
    from gecoz_b200 import synth
    seq = synth.chromosome_shaped(40_000, 5)
    text = synth.block_of([seq])
    r = O.build_block(text, 32, want_sa=True)
    g = O.GSSA(r["gcz_body"], len(text), r["gcx_body"])
    isa = np.zeros(len(text), np.int64)
    isa[r["sa"]] = np.arange(len(text))
    for p in (0, 32, 64, 4096, 39_968, 40_000):                  # sampled positions (40 000 is the separator)
        assert g.index_find(p) == isa[p]
    assert g.index_find(33) == -(1 << 31)                        # Integer.MIN_VALUE for an unsampled position
    for start, cap in ((0, 40_000), (0, 50_000), (0, 1), (31, 2), (39_990, 100), (12_345, 6_789)):
        assert np.array_equal(g.extract(0, start, cap), seq[start:start + cap])
    with pytest.raises(IndexError):
        g.extract(1, 0, 5)


@pytest.mark.parametrize("seed", range(6))
def test_select_and_inverse_lookups_are_exact(seed):
    """RankedWTNode.findOne/findZero (:130-205) return the exact select; IndexWaveletTree.find (:152-165) inverts get;
    GSSAIndex.find (:184-187) inverts locate at the sampled positions — on random small inputs."""
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(1, 200_000))
    bits = (rng.random(n) < rng.choice([0.02, 0.5, 0.97])).astype(np.uint8)
    buf = O.ranked_write(bits)
    ones, zeros = np.flatnonzero(bits), np.flatnonzero(bits == 0)
    for k in rng.integers(1, max(2, len(ones) + 1), 40):
        if k <= len(ones):
            assert O.ranked_find_one(buf, n, int(k)) == ones[k - 1]
    for k in rng.integers(1, max(2, len(zeros) + 1), 40):
        if k <= len(zeros):
            assert O.ranked_find_zero(buf, n, int(k)) == zeros[k - 1]
    m = int(rng.integers(1, 5000))
    vals = rng.permutation(m).astype(np.int32)
    iwt = O.iwt_write(vals)
    for p in rng.integers(0, m, 60):
        assert O.iwt_find(iwt, m, int(vals[p])) == p
    from gecoz_b200 import synth
    text = synth.block_of([synth.iid_acgtn(int(rng.integers(40, 3000)), seed)])
    rate = int(rng.choice([2, 8, 32]))
    r = O.build_block(text, rate, want_sa=True)
    g = O.GSSA(r["gcz_body"], len(text), r["gcx_body"])
    isa = np.zeros(len(text), np.int64)
    isa[r["sa"]] = np.arange(len(text))
    if g.sampling_factor != rate.bit_length() - 1:
        # the factor is not stored: readers take the smallest one whose index fits (algo/ssa/GSSAIndex.java:62-67), and
        # for a tiny text two factors can give the same size — the reference then mis-reads its own index
        return
    for p in range(0, len(text), rate):
        assert g.index_find(p) == isa[p]
    start = int(rng.integers(0, len(text) - 1))
    assert np.array_equal(g.extract(0, start, 10_000), text[start:-1])
