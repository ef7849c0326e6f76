"""GSSA — the generalized succinct suffix array of one block, resident on a GPU.

Mirror of algo/ssa/GSSA.java (/root/reference/java/nova-algo/src/main/java/es/elixir/bsc/ngs/nova/algo/ssa/GSSA.java):
`find` :160-185, `count` :136-148, `getLength` :67-88 keep their single-pattern signatures; the `*_batch`
forms are what the GPU is for (SimpleGFFGenerator-style callers collect patterns and call once).
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _native as N


def pack_patterns(patterns: Sequence[bytes]) -> tuple[np.ndarray, np.ndarray]:
    """Concatenate patterns into (bytes, offsets[n+1]) — the layout gcz_count_batch / gcz_find_batch take."""
    off = np.zeros(len(patterns) + 1, dtype=np.int64)
    np.cumsum([len(p) for p in patterns], out=off[1:])
    data = np.frombuffer(b"".join(bytes(p) for p in patterns), dtype=np.uint8).copy() if len(patterns) else np.zeros(0, np.uint8)
    if len(data) == 0:
        data = np.zeros(1, np.uint8)
    return data, off


class GSSA:
    def __init__(self, handle, device: int, headers: Sequence[str] | None = None):
        self._h = handle
        self.device = device
        self.headers = list(headers) if headers is not None else None
        L = N.lib()
        n, sf, ns = C.c_int64(), C.c_int32(), C.c_int32()
        N.check(L.gcz_text_length(self._h, C.byref(n)))
        N.check(L.gcz_sampling_factor(self._h, C.byref(sf)))
        N.check(L.gcz_num_strings(self._h, C.byref(ns)))
        self.length, self.sampling_factor, self.n_strings = n.value, sf.value, ns.value
        self.e = np.zeros(max(self.n_strings, 1), dtype=np.int64)
        N.check(L.gcz_string_ends(self._h, N.ptr(self.e)))
        self.e = self.e[:self.n_strings]
        self.c = np.zeros(256, dtype=np.int64)
        N.check(L.gcz_c_array(self._h, N.ptr(self.c)))

    @classmethod
    def open(cls, device: int, gcz_body, text_len: int, gcx_body, headers=None) -> "GSSA":
        """GecozFileReader.read's tail: bodies (host or device memory) -> resident index."""
        h = C.c_void_p()
        nb = gcz_body.numel() if hasattr(gcz_body, "numel") else len(gcz_body)
        nx = gcx_body.numel() if hasattr(gcx_body, "numel") else len(gcx_body)
        if isinstance(gcz_body, np.ndarray) and not gcz_body.flags["C_CONTIGUOUS"]:
            gcz_body = np.ascontiguousarray(gcz_body)
        if isinstance(gcx_body, np.ndarray) and not gcx_body.flags["C_CONTIGUOUS"]:
            gcx_body = np.ascontiguousarray(gcx_body)
        N.check(N.lib().gcz_open_block(device, N.ptr(gcz_body), nb, text_len, N.ptr(gcx_body), nx, C.byref(h)))
        return cls(h, device, headers)

    def close(self):
        if self._h:
            N.lib().gcz_close_block(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- GSSA API ------------------------------------------------------------------------------------
    def getLength(self, nstr: int | None = None) -> int:            # :67-88
        if nstr is None:
            return self.length
        if nstr < 0 or nstr >= self.n_strings:
            raise IndexError(f"String index {nstr} is out of bound")
        return int(self.e[nstr]) if nstr == 0 else int(self.e[nstr] - self.e[nstr - 1] - 1)

    def count_batch(self, patterns=None, packed=None, out=None):
        """Backward-search intervals: (sp, ep) int64 arrays; occurrences = max(0, ep - sp + 1)."""
        data, off = packed if packed is not None else pack_patterns(patterns)
        n = (off.numel() if hasattr(off, "numel") else len(off)) - 1
        if out is None:
            sp, ep = np.zeros(n, np.int64), np.zeros(n, np.int64)
        else:
            sp, ep = out
        N.check(N.lib().gcz_count_batch(self._h, N.ptr(data), N.ptr(off), n, N.ptr(sp), N.ptr(ep)))
        return sp, ep

    def locate_rows(self, rows) -> np.ndarray:
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        out = np.zeros(len(rows), np.int64)
        N.check(N.lib().gcz_locate_rows(self._h, N.ptr(rows), len(rows), N.ptr(out)))
        return out

    def find_batch_raw(self, patterns=None, packed=None):
        """GSSA.find for a batch, as arrays: (per_string_counts [n, n_strings], positions, pos_off[n + 1]).
        positions holds, pattern after pattern and string after string, the ascending 0-based positions
        relative to the string start."""
        data, off = packed if packed is not None else pack_patterns(patterns)
        data, off = np.ascontiguousarray(data, dtype=np.uint8), np.ascontiguousarray(off, dtype=np.int64)
        n, ns = len(off) - 1, self.n_strings
        per = np.zeros(max(n * ns, 1), dtype=np.int64)
        ppos, poff = C.c_void_p(), C.c_void_p()
        N.check(N.lib().gcz_find_batch(self._h, N.ptr(data), N.ptr(off), n, N.ptr(per), C.byref(ppos), C.byref(poff)))
        try:
            offs = np.ctypeslib.as_array(C.cast(poff, C.POINTER(C.c_int64)), shape=(n + 1,)).copy()
            total = int(offs[-1])
            pos = (np.ctypeslib.as_array(C.cast(ppos, C.POINTER(C.c_int64)), shape=(max(total, 1),)).copy()[:total])
        finally:
            N.lib().gcz_free(ppos)
            N.lib().gcz_free(poff)
        per = per[:n * ns].reshape(n, ns) if ns else per[:0].reshape(n, 0)
        return per, pos, offs

    def find_batch(self, patterns: Sequence[bytes]):
        """GSSA.find for every pattern: list of (None | list per string of (None | int64 array))."""
        data, off = pack_patterns(patterns)
        n, ns = len(patterns), self.n_strings
        per, pos, offs = self.find_batch_raw(packed=(data, off))
        sp, ep = self.count_batch(packed=(data, off))
        res = []
        for i in range(n):
            if ep[i] < sp[i]:
                res.append(None)                      # search() returned an empty array -> find returns null
                continue
            o, one = int(offs[i]), []
            for s in range(ns):
                k = int(per[i, s])
                one.append(pos[o:o + k] if k > 0 else None)
                o += k
            res.append(one)
        return res

    def extract(self, nstr: int, start: int = 0, length: int | None = None, out=None) -> np.ndarray:
        """GSSA.extract(buf, nstr, from)  :90-126: up to `length` symbols of string nstr from `start` (never past the
        string's end).  Returns the bytes written (a view of `out` when one is given)."""
        if nstr < 0 or nstr >= self.n_strings:
            raise IndexError(f"String index {nstr} is out of bound")
        cap = self.getLength(nstr) - start if length is None else length
        cap = max(int(cap), 0)
        buf = np.zeros(max(cap, 1), dtype=np.uint8) if out is None else out
        w = C.c_int64()
        N.check(N.lib().gcz_extract(self._h, nstr, int(start), N.ptr(buf), cap, C.byref(w)))
        return buf[:w.value]

    def find(self, pattern: bytes):                                    # :160-185
        return self.find_batch([pattern])[0]

    def count(self, pattern: bytes):                                   # :136-148
        sa = self.find(pattern)
        if sa is None or len(sa) == 0:
            return None
        return [len(x) if x is not None and len(x) > 0 else 0 for x in sa]


# ---- a batch against every block of a file (include/gcz.h: gcz_count_multi / gcz_find_multi) -------------------------------
def _handles(gssas):
    arr = (C.c_void_p * len(gssas))(*[g._h for g in gssas])
    return arr, len(gssas)


def _count(n_or_off):
    return (n_or_off.numel() if hasattr(n_or_off, "numel") else len(n_or_off)) - 1


def count_totals(gssas: Sequence[GSSA], data, off, out=None):
    """Occurrences of every pattern summed over the blocks (GecoMatch's "total found", tools/GecoMatch.java:114-131): the
    batch (host or device arrays) is uploaded once and searched against every block on the device.  `out`: int64[n]
    (numpy, pinned or CUDA tensor); a numpy array is returned when it is omitted."""
    n = _count(off)
    if out is None:
        out = np.zeros(n, np.int64)
    arr, k = _handles(gssas)
    N.check(N.lib().gcz_count_multi(arr, k, N.ptr(data), N.ptr(off), n, N.ptr(out)))
    return out


def count_stats(gssas: Sequence[GSSA], data, off) -> dict:
    """Rank sectors read / backward-search steps / reference rank calls of a batch (gcz_count_stats; measurement only)."""
    st = N.QueryStats()
    arr, k = _handles(gssas)
    N.check(N.lib().gcz_count_stats(arr, k, N.ptr(data), N.ptr(off), _count(off), C.byref(st)))
    return st.as_dict()


def last_query_stats() -> dict:
    st = N.QueryStats()
    N.check(N.lib().gcz_last_query_stats(C.byref(st)))
    return st.as_dict()


def find_multi(gssas: Sequence[GSSA], data, off, copy: bool = True):
    """GSSA.find of a batch against every block: (block_off[k + 1], pattern, string, position) — the hits of block b are
    rows block_off[b] .. block_off[b + 1], sorted by (pattern, string, position)."""
    hits = N.Hits()
    arr, k = _handles(gssas)
    N.check(N.lib().gcz_find_multi(arr, k, N.ptr(data), N.ptr(off), _count(off), C.byref(hits)))
    try:
        n = int(hits.n_hits)
        block_off = np.ctypeslib.as_array(hits.block_off, shape=(k + 1,)).copy()
        if not copy:
            return block_off, n
        if n == 0:
            return block_off, np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int64)
        return (block_off, np.ctypeslib.as_array(hits.pattern, shape=(n,)).copy(), np.ctypeslib.as_array(hits.string, shape=(n,)).copy(),
                np.ctypeslib.as_array(hits.position, shape=(n,)).copy())
    finally:
        N.lib().gcz_hits_free(C.byref(hits))


def find_total(gssas: Sequence[GSSA], data, off) -> int:
    """Number of hits of a batch over all blocks (the arrays are produced and dropped: what a caller that streams them out pays)."""
    return int(find_multi(gssas, data, off, copy=False)[1])
