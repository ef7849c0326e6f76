"""BASELINE.json full sizes on the GPU.

cfg1 (16 Mbp) is still within the oracle's reach: full byte parity.  cfg2 (248 956 422 bp, chr1-shaped) is
checked through size-independent properties: the suffix array is a sorted permutation (sampled adjacent
pairs compared on the text), BWT == text[SA-1], the index answers locate(row) == SA[row], every text-sampled
pattern is found at positions where the text really holds it (build -> open -> count -> locate round trip),
and two builds are byte-identical.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gecoz_b200 as g
    g.lib()
    return g


def _build(G, text, want_sa=True):
    shape = G.shape_from_counts(G.symbol_counts(text))
    gcz = np.zeros(shape.size, np.uint8)
    gcx = np.zeros(G.index_size(len(text), 5), np.uint8)
    sa = np.zeros(len(text), np.int32) if want_sa else None
    bwt = np.zeros(len(text), np.uint8) if want_sa else None
    t = G.build_block(0, text, len(text), 32, shape, gcz, gcx, sa, bwt)
    return shape, gcz, gcx, sa, bwt, t


def test_cfg1_full_parity(G):
    from gecoz_b200 import synth
    from oracle import gcz_oracle as O
    text = synth.cfg1_text()
    assert len(text) == 16_000_001
    shape, gcz, gcx, sa, bwt, t = _build(G, text)
    ref = O.build_block(text, 32, want_sa=True, want_bwt=True, threads=2)
    assert np.array_equal(sa, ref["sa"]) and np.array_equal(bwt, ref["bwt"])
    assert np.array_equal(gcz, ref["gcz_body"]) and np.array_equal(gcx, ref["gcx_body"])
    assert len(gcx) == 3_289_370
    # 10k random 15-mers: half from the text, half i.i.d. (config 1 of BASELINE.json)
    g = G.GSSA.open(0, gcz, len(text), gcx)
    og = O.GSSA(ref["gcz_body"], len(text), ref["gcx_body"])
    data, off = synth.patterns(text, 10_000, 15, 15, seed=2)
    sp, ep = g.count_batch(packed=(data, off))
    esp, eep, _ = og.search_batch(data, off)
    assert np.array_equal(sp, esp) and np.array_equal(ep, eep)
    assert 4_900 < int((ep >= sp).sum()) < 5_300
    g.close()


def _suffix_less_equal(text, a, b):
    """text[a:] <= text[b:] under unsigned bytes with 'a proper prefix sorts first'."""
    n = len(text)
    step = 1 << 12
    while True:
        la, lb = min(step, n - a), min(step, n - b)
        l = min(la, lb)
        x, y = text[a:a + l], text[b:b + l]
        d = np.flatnonzero(x != y)
        if len(d):
            return x[d[0]] < y[d[0]]
        if l < step:                       # one of them ended
            return (n - a) <= (n - b)
        a += l
        b += l
        step = min(step * 4, 1 << 24)


def test_cfg2_full_properties(G):
    import torch
    from gecoz_b200 import synth
    text = synth.cfg2_text()
    n = len(text)
    assert n == 248_956_423
    shape, gcz, gcx, sa, bwt, t = _build(G, text)
    assert len(gcx) == 55_197_282 and len(gcz) == shape.size
    # permutation
    dsa = torch.from_numpy(sa).cuda()
    srt, _ = torch.sort(dsa)
    assert bool((srt == torch.arange(n, dtype=torch.int32, device="cuda")).all())
    del srt
    # BWT == text[SA - 1] (text[n-1] for SA == 0)
    dtext = torch.from_numpy(text).cuda()
    idx = dsa.long() - 1
    idx[idx < 0] = n - 1
    assert bool((dtext[idx] == torch.from_numpy(bwt).cuda()).all())
    del idx, dtext, dsa
    torch.cuda.empty_cache()
    # sortedness on sampled adjacent rows, including rows deep inside the 18 Mbp N run
    rng = np.random.default_rng(0)
    rows = np.concatenate([rng.integers(0, n - 1, 1500), np.arange(0, 64), np.arange(n - 65, n - 1)])
    first_n = int((text < ord("N")).sum())                       # rows of suffixes starting with N begin here
    rows = np.concatenate([rows, first_n + rng.integers(0, 18_000_000, 200)])
    for r in rows:
        assert _suffix_less_equal(text, int(sa[r]), int(sa[r + 1])), r
    assert sa[0] == n - 1
    # round trip through the index
    g = G.GSSA.open(0, gcz, n, gcx)
    assert g.n_strings == 1 and g.e.tolist() == [n - 1] and g.sampling_factor == 5
    cnt = np.bincount(text, minlength=256)
    assert np.array_equal(g.c, np.cumsum(cnt) - cnt)
    lr = rng.integers(0, n, 100_000)
    assert np.array_equal(g.locate_rows(lr), sa[lr].astype(np.int64))
    data, off = synth.patterns(text, 20_000, 15, 100, seed=5)
    sp, ep = g.count_batch(packed=(data, off))
    hits = np.flatnonzero(ep >= sp)
    assert len(hits) > 9_500
    for q in hits[:300]:
        pat = data[off[q]:off[q + 1]]
        pos = g.locate_rows(np.arange(sp[q], ep[q] + 1))
        for p in pos[:4]:
            assert np.array_equal(text[p:p + len(pat)], pat)
        assert np.array_equal(np.sort(sa[sp[q]:ep[q] + 1]), np.sort(pos))
    g.close()
    # determinism: a second build writes the same bytes
    shape2, gcz2, gcx2, _, _, _ = _build(G, text, want_sa=False)
    assert np.array_equal(gcz, gcz2) and np.array_equal(gcx, gcx2)


# ---- byte parity at the benchmarked sizes: sha256 against the oracle's digests --------------------------------------
# tests/golden/full_size_digests.json is written by tools/make_digests.py from the CPU oracle alone (5 minutes on 8
# cores); nothing here runs the oracle, so the comparison costs one build and one hash per block on the GPU box.
import hashlib
import json
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

GOLD = json.loads((Path(__file__).parent / "golden" / "full_size_digests.json").read_text())


def _sha(a) -> str:
    return hashlib.sha256(memoryview(np.ascontiguousarray(a)).cast("B")).hexdigest()


def test_cfg2_digests_equal_the_oracles(G):
    from gecoz_b200 import synth
    text = synth.cfg2_text()
    g2 = GOLD["cfg2"]
    assert len(text) == g2["n"] and _sha(text) == g2["text"], "the generator no longer produces the text the digests were made from"
    shape, gcz, gcx, sa, bwt, _ = _build(G, text)
    assert (len(gcz), len(gcx)) == (g2["gcz_body_len"], g2["gcx_body_len"])
    assert _sha(gcz) == g2["gcz_body"], ".gcz body of the full cfg2 block differs from the oracle's"
    assert _sha(gcx) == g2["gcx_body"], ".gcx body of the full cfg2 block differs from the oracle's"
    assert _sha(sa) == g2["sa"], "suffix array of the full cfg2 block differs from the oracle's"
    assert _sha(bwt) == g2["bwt"], "BWT of the full cfg2 block differs from the oracle's"


@pytest.fixture(scope="module")
def genome():
    """The hg38-shaped genome as the product blocks it (GecoIndex merge) and its synthetic block texts (made on demand)."""
    from gecoz_b200 import synth
    from gecoz_b200.geco_index import FastaSequence, merge_blocks
    seqs = [FastaSequence(h, ln, None, i) for i, (h, ln) in enumerate(zip(synth.HG38_NAMES, synth.HG38_LENGTHS))]
    plan = [([s.header for s in b.sequences], [s.id for s in b.sequences], int(b.size)) for b in merge_blocks(seqs)]

    def text_of(b):
        return synth.block_of([synth.chromosome_shaped(synth.HG38_LENGTHS[i], 4 + i) for i in plan[b][1]])
    return plan, text_of


def test_cfg3_block_plan_equals_the_oracles(genome):
    plan, _ = genome
    gold = GOLD["cfg3"]["blocks"]
    assert [p[0] for p in plan] == [b["headers"] for b in gold]
    assert [p[2] for p in plan] == [b["n"] for b in gold]
    assert sum(p[2] for p in plan) == GOLD["cfg3"]["symbols"] == 3_088_286_426


def test_cfg3_every_block_equals_the_oracles_digests(G, genome):
    """All 18 blocks of the 3.1 Gbp genome: .gcz and .gcx bodies (and SA + BWT on the merged blocks) byte for byte, and the two
    whole files (headers + bodies in file order) against the digests of the files the oracle writes."""
    from gecoz_b200.gecoz_file import GecozRefBlockHeader, GecozSSABlockHeader
    plan, text_of = genome
    gold = GOLD["cfg3"]
    hz, hx = hashlib.sha256(), hashlib.sha256()
    with ThreadPoolExecutor(3) as pool:                       # synthesis of the next blocks overlaps build + hashing
        texts = [pool.submit(text_of, b) for b in range(len(plan))]
        for b, (headers, ids, n) in enumerate(plan):
            text = texts[b].result()
            texts[b] = None
            gb = gold["blocks"][b]
            assert len(text) == n and _sha(text) == gb["text"], f"block {b}: synthetic text differs from the one the digests were made from"
            merged = len(ids) > 1
            shape, gcz, gcx, sa, bwt, _ = _build(G, text, want_sa=merged)
            assert _sha(gcz) == gb["gcz_body"], f"block {b} {headers}: .gcz body differs from the oracle's"
            assert _sha(gcx) == gb["gcx_body"], f"block {b} {headers}: .gcx body differs from the oracle's"
            if merged:
                assert _sha(sa) == gb["sa"] and _sha(bwt) == gb["bwt"], f"block {b} {headers}: SA / BWT differ from the oracle's"
            hz.update(GecozRefBlockHeader(headers, GecozRefBlockHeader.block_header_length(headers) + len(gcz), n).to_bytes())
            hz.update(memoryview(gcz))
            hx.update(GecozSSABlockHeader(headers, len(gcx)).to_bytes())
            hx.update(memoryview(gcx))
            del text, gcz, gcx, sa, bwt
    assert hz.hexdigest() == gold["gcz_file"] and hx.hexdigest() == gold["gcx_file"]


@pytest.mark.parametrize("first", ["chr13", "chr15", "chr11"])
def test_cfg4_cfg5_results_equal_the_oracles_digests(G, genome, first):
    """count (cfg4) and find (cfg5) at full size on the merged blocks chr13+chr14 and chr15+chr22+chr21+chrM — where per-string
    results are more than intervals — and on the chr11 block of the `-s chr11` filter: 100 000 patterns of 15..100 symbols."""
    from gecoz_b200 import synth
    import gecoz_b200 as GG
    plan, text_of = genome
    b = [p[0][0] for p in plan].index(first)
    q = GOLD["cfg3"]["blocks"][b]["queries"]
    text = text_of(b)
    shape, gcz, gcx, _, _, _ = _build(G, text, want_sa=False)
    g = G.GSSA.open(0, gcz, len(text), gcx, plan[b][0])
    assert g.n_strings == q["n_strings"] and g.e.tolist() == q["string_ends"]
    data, off = synth.patterns(text, q["patterns"], 15, 100, seed=q["seed"])
    assert _sha(data) == q["pattern_bytes"]
    sp, ep = g.count_batch(packed=(data, off))
    assert _sha(sp) == q["sp"] and _sha(ep) == q["ep"] and int((ep >= sp).sum()) == q["found"]
    # the rank calls the reference's loop makes for this batch, counted by the kernel, equal the instrumented oracle's
    st = GG.count_stats([g], data, off)
    assert st["reference_rank_calls"] == q["rank_calls"]
    assert 0 < st["rank_sectors"] <= st["reference_rank_calls"]
    # totals over "all blocks" of a one-block index
    tot = GG.count_totals([g], data, off)
    assert np.array_equal(tot, np.maximum(ep - sp + 1, 0))
    k = q["find_patterns"]
    per, pos, poff = g.find_batch_raw(packed=(data[:off[k]], off[:k + 1]))
    assert _sha(per) == q["per_string_counts"] and _sha(pos) == q["positions"] and len(pos) == q["occurrences"]
    g.close()
