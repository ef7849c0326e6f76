"""Fuzz of the suffix sorter's algorithm (tests/sorter_model.py) against the oracle's suffix array: forced small
key widths, run-heavy and tie-heavy texts, several strings per block."""
import numpy as np
import pytest

from oracle import gcz_oracle as O
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
from sorter_model import suffix_array_model  # noqa: E402


def _texts(rng, count):
    for _ in range(count):
        kind = rng.integers(0, 5)
        n = int(rng.integers(1, 120))
        if kind == 0:                                     # random over a tiny alphabet
            t = rng.integers(1, int(rng.integers(2, 5)), n).astype(np.uint8) + 64
        elif kind == 1:                                   # runs of random lengths
            lens = rng.integers(1, 12, max(1, n // 4))
            t = np.repeat(rng.integers(65, 69, len(lens)).astype(np.uint8), lens)
        elif kind == 2:                                   # equal-length runs, varying terminators
            parts = []
            for i in range(max(1, n // 8)):
                parts += [np.full(int(rng.integers(3, 7)), 78, np.uint8), rng.integers(65, 69, int(rng.integers(1, 3))).astype(np.uint8)]
            t = np.concatenate(parts)
        elif kind == 3:                                   # periodic
            p = rng.integers(65, 68, int(rng.integers(1, 4))).astype(np.uint8)
            t = np.tile(p, n // len(p) + 1)[:n]
        else:                                             # several strings, empty ones included
            parts = []
            for i in range(int(rng.integers(1, 6))):
                parts += [rng.integers(65, 67, int(rng.integers(0, 15))).astype(np.uint8), np.zeros(1, np.uint8)]
            t = np.concatenate(parts)[:-1]
        yield np.concatenate([t, np.zeros(1, np.uint8)])


@pytest.mark.parametrize("seed", range(4))
def test_algorithm_matches_the_oracle(seed):
    rng = np.random.default_rng(seed)
    for text in _texts(rng, 150):
        exp = O.suffix_array(text)
        for k in (1, 2, 3, 5, None):
            got = suffix_array_model(text, k)
            assert np.array_equal(got, exp), (text.tobytes(), k)
        assert np.array_equal(suffix_array_model(text, 2, use_runs=False), exp)     # plain doubling (mark-buffer overflow path)
