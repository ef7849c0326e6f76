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
