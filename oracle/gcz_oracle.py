"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_build/libgcz_oracle.so (the CPU restatement of gecoz's FM-index
path, see oracle/orc_internal.h) plus the file-level composition the reference performs in
fmt/GecozFileWriter.java:124-159 and tools/GecoIndex.java:51-146.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package gecoz_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "_build" / "libgcz_oracle.so"


def build(force: bool = False) -> Path:
    """Compile the C restatement (gcc, seconds).  Building the checker is not using it."""
    if force or not _SO.exists() or any(
        p.stat().st_mtime > _SO.stat().st_mtime for p in list(_DIR.glob("*.c")) + list(_DIR.glob("*.h"))
    ):
        subprocess.run(["make", "-C", str(_DIR)], check=True, stdout=subprocess.DEVNULL)
    return _SO


class Shape(C.Structure):
    _fields_ = [
        ("bit_lengths", C.c_int8 * 256),
        ("table", C.c_int16 * 256),
        ("node_bits", C.c_int32 * 256),
        ("length", C.c_int64),
        ("size", C.c_int64),
        ("table_bytes", C.c_int64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_SO))
        P = C.c_void_p
        i32, i64 = C.c_int32, C.c_int64
        sig = {
            "orc_bitbuffer_write": (i64, [P, P, i32, P, i64]),
            "orc_bitbuffer_write_read": (None, [P, P, i32, i64, P, i32, P]),
            "orc_ranked_bytes": (i64, [i64]),
            "orc_ranked_write": (i64, [P, i64, P]),
            "orc_ranked_get": (i32, [P, i64, i64]),
            "orc_ranked_count": (i64, [P, i64, i64]),
            "orc_ranked_find_one": (i64, [P, i64, i64]),
            "orc_ranked_find_zero": (i64, [P, i64, i64]),
            "orc_deflate_encode_table": (i32, [P, i32, i32, P, P]),
            "orc_lookup_get_symbol": (i32, [P, i32, i32]),
            "orc_lookup_get_symbol_nbits": (i32, [P, i32, i32, i32]),
            "orc_deflate_stream_roundtrip": (i64, [P, i64, i64]),
            "orc_deflate_lengths_bits": (i32, [P, i32]),
            "orc_shape_from_counts": (i32, [P, C.POINTER(Shape)]),
            "orc_shape_write": (i64, [C.POINTER(Shape), P, i64]),
            "orc_suffix_array": (i32, [P, i64, P]),
            "orc_suffix_array_naive": (i32, [P, i64, P]),
            "orc_index_size": (i64, [i64, i32]),
            "orc_build_block": (i32, [P, i64, i32, P, i64, P, i64, P, P, i32]),
            "orc_hswt_write": (i32, [C.POINTER(Shape), P, P, i64, P, i64]),
            "orc_gssa_index_write": (i32, [P, i64, i32, P, i64]),
            "orc_iwt_write": (i32, [P, i64, P, i64]),
            "orc_iwt_get": (i64, [P, i64, i64]),
            "orc_iwt_find": (i64, [P, i64, i64]),
            "orc_open": (P, [P, i64, i64, P, i64]),
            "orc_close": (None, [P]),
            "orc_sampling_factor": (i32, [P]),
            "orc_num_strings": (i32, [P]),
            "orc_string_ends": (None, [P, P]),
            "orc_c_array": (None, [P, P]),
            "orc_num_nodes": (i32, [P]),
            "orc_node_info": (None, [P, P, P, P]),
            "orc_occ": (i64, [P, i32, i64]),
            "orc_get_rs": (i64, [P, i64]),
            "orc_search": (i64, [P, P, i64, P, P]),
            "orc_locate": (i64, [P, i64]),
            "orc_index_find": (i64, [P, i64]),
            "orc_extract": (i64, [P, C.c_int32, i64, P, i64]),
            "orc_find": (i64, [P, P, i64, P, P, i64]),
            "orc_search_batch": (i64, [P, P, P, i64, P, P]),
            "orc_rank_calls": (C.c_uint64, []),
            "orc_rank_calls_reset": (None, []),
            "orc_header_hash": (i64, [P, i32]),
            "orc_ref_header_len": (i32, [P, i32]),
            "orc_ref_header_write": (i32, [P, i32, i64, i64, P]),
            "orc_ssa_header_write": (i32, [P, i32, i64, P]),
            "orc_merge_blocks": (i32, [P, P, i32, P, P]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _u8(x) -> np.ndarray:
    if isinstance(x, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(x), dtype=np.uint8).copy()
    return np.ascontiguousarray(x, dtype=np.uint8)


def _cstrs(headers):
    arr = (C.c_char_p * len(headers))(*[h.encode("ascii") for h in headers])
    return arr


# ---- bit stream / ranked vector ----------------------------------------------------------
def bitbuffer_write(vals, nbits, cap: int) -> tuple[bytes, int]:
    v = np.asarray(vals, dtype=np.int64)
    nb = np.asarray(nbits, dtype=np.int32)
    out = np.zeros(cap + 8, dtype=np.uint8)
    pos = lib().orc_bitbuffer_write(_p(v), _p(nb), len(v), _p(out), cap)
    return out[:cap].tobytes(), int(pos)


def bitbuffer_write_read(vals, nbits, cap: int, read_nbits):
    v = np.asarray(vals, dtype=np.int64)
    nb = np.asarray(nbits, dtype=np.int32)
    rn = np.asarray(read_nbits, dtype=np.int32)
    out = np.zeros(len(rn), dtype=np.int64)
    lib().orc_bitbuffer_write_read(_p(v), _p(nb), len(v), cap, _p(rn), len(rn), _p(out))
    return out


def ranked_bytes(length: int) -> int:
    return int(lib().orc_ranked_bytes(length))


def ranked_write(bits) -> np.ndarray:
    b = _u8(bits)
    out = np.zeros(ranked_bytes(len(b)) + 8, dtype=np.uint8)
    n = lib().orc_ranked_write(_p(b), len(b), _p(out))
    return out[:n].copy()


def ranked_count(buf: np.ndarray, length: int, idx: int) -> int:
    return int(lib().orc_ranked_count(_p(buf), length, idx))


def ranked_get(buf: np.ndarray, length: int, idx: int) -> int:
    return int(lib().orc_ranked_get(_p(buf), length, idx))


def ranked_find_one(buf, length, n):
    return int(lib().orc_ranked_find_one(_p(buf), length, n))


def ranked_find_zero(buf, length, n):
    return int(lib().orc_ranked_find_zero(_p(buf), length, n))


# ---- tables ---------------------------------------------------------------------------------
def deflate_encode_table(counts, max_bits: int = 15):
    c = np.ascontiguousarray(counts, dtype=np.int64)
    bl = np.zeros(len(c), dtype=np.int8)
    tb = np.zeros(len(c), dtype=np.int16)
    rc = lib().orc_deflate_encode_table(_p(c), len(c), max_bits, _p(bl), _p(tb))
    if rc != 0:
        raise RuntimeError(f"orc_deflate_encode_table rc={rc}")
    return bl, tb


def lookup_get_symbol(bit_lengths, code: int, nbits: int | None = None) -> int:
    bl = np.ascontiguousarray(bit_lengths, dtype=np.int8)
    if nbits is None:
        return int(lib().orc_lookup_get_symbol(_p(bl), len(bl), code))
    return int(lib().orc_lookup_get_symbol_nbits(_p(bl), len(bl), code, nbits))


def deflate_stream_roundtrip(data, cap: int) -> int:
    d = _u8(data)
    return int(lib().orc_deflate_stream_roundtrip(_p(d), len(d), cap))


def shape_from_counts(counts) -> Shape:
    c = np.ascontiguousarray(counts, dtype=np.int64)
    assert len(c) == 256
    s = Shape()
    rc = lib().orc_shape_from_counts(_p(c), C.byref(s))
    if rc != 0:
        raise RuntimeError(f"orc_shape_from_counts rc={rc}")
    return s


def shape_write(s: Shape) -> bytes:
    out = np.zeros(int(s.table_bytes) + 16, dtype=np.uint8)
    n = lib().orc_shape_write(C.byref(s), _p(out), len(out))
    return out[:n].tobytes()


# ---- suffix array / block build ---------------------------------------------------------------
def suffix_array(text, naive: bool = False) -> np.ndarray:
    t = _u8(text)
    sa = np.zeros(len(t), dtype=np.int32)
    f = lib().orc_suffix_array_naive if naive else lib().orc_suffix_array
    rc = f(_p(t), len(t), _p(sa))
    if rc != 0:
        raise RuntimeError(f"suffix_array rc={rc}")
    return sa


def index_size(n: int, sampling_factor: int) -> int:
    return int(lib().orc_index_size(n, sampling_factor))


def build_block(text, sampling_rate: int = 32, want_sa: bool = False, want_bwt: bool = False, threads: int = 1):
    """BlockWriter.run on one generalized string: returns dict(gcz_body, gcx_body[, sa, bwt])."""
    t = _u8(text)
    n = len(t)
    counts = np.bincount(t, minlength=256).astype(np.int64)
    s = shape_from_counts(counts)
    gcz = np.zeros(int(s.size), dtype=np.uint8)
    gcx = np.zeros(index_size(n, sampling_rate.bit_length() - 1), dtype=np.uint8)
    sa = np.zeros(n, dtype=np.int32) if want_sa else None
    bwt = np.zeros(n, dtype=np.uint8) if want_bwt else None
    rc = lib().orc_build_block(_p(t), n, sampling_rate, _p(gcz), len(gcz), _p(gcx), len(gcx),
                               _p(sa) if sa is not None else None, _p(bwt) if bwt is not None else None, threads)
    if rc != 0:
        raise RuntimeError(f"orc_build_block rc={rc}")
    res = {"gcz_body": gcz, "gcx_body": gcx, "shape": s}
    if want_sa:
        res["sa"] = sa
    if want_bwt:
        res["bwt"] = bwt
    return res


def iwt_write(vals) -> np.ndarray:
    v = np.ascontiguousarray(vals, dtype=np.int32)
    m = len(v)
    size = ranked_bytes(m) * int(m).bit_length()
    out = np.zeros(size + 8, dtype=np.uint8)
    rc = lib().orc_iwt_write(_p(v), m, _p(out), size)
    if rc != 0:
        raise RuntimeError(f"orc_iwt_write rc={rc}")
    return out[:size].copy()


def iwt_get(buf: np.ndarray, m: int, pos: int) -> int:
    return int(lib().orc_iwt_get(_p(buf), m, pos))


def iwt_find(buf: np.ndarray, m: int, idx: int) -> int:
    return int(lib().orc_iwt_find(_p(buf), m, idx))


def gssa_index_write(sa, sampling_rate: int = 32) -> np.ndarray:
    s = np.ascontiguousarray(sa, dtype=np.int32)
    size = index_size(len(s), sampling_rate.bit_length() - 1)
    out = np.zeros(size + 8, dtype=np.uint8)
    rc = lib().orc_gssa_index_write(_p(s), len(s), sampling_rate, _p(out), size)
    if rc != 0:
        raise RuntimeError(f"orc_gssa_index_write rc={rc}")
    return out[:size].copy()


# ---- reader / GSSA --------------------------------------------------------------------------------
class GSSA:
    """algo/ssa/GSSA.java over one block's bodies (what GecozFileReader.read returns)."""

    def __init__(self, gcz_body, text_len: int, gcx_body):
        self._gcz = _u8(gcz_body)
        self._gcx = _u8(gcx_body)
        # 8 bytes of slack: the literal reader peeks whole longs
        self._gcz_pad = np.concatenate([self._gcz, np.zeros(8, np.uint8)])
        self._gcx_pad = np.concatenate([self._gcx, np.zeros(8, np.uint8)])
        self.h = lib().orc_open(_p(self._gcz_pad), len(self._gcz), text_len, _p(self._gcx_pad), len(self._gcx))
        if not self.h:
            raise RuntimeError("orc_open failed")
        self.text_len = text_len
        self.n_strings = int(lib().orc_num_strings(self.h))

    def close(self):
        if self.h:
            lib().orc_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def sampling_factor(self) -> int:
        return int(lib().orc_sampling_factor(self.h))

    def string_ends(self) -> np.ndarray:
        e = np.zeros(self.n_strings, dtype=np.int64)
        lib().orc_string_ends(self.h, _p(e))
        return e

    def c_array(self) -> np.ndarray:
        c = np.zeros(256, dtype=np.int64)
        lib().orc_c_array(self.h, _p(c))
        return c

    def nodes(self):
        k = int(lib().orc_num_nodes(self.h))
        names = np.zeros(k, np.int32)
        lens = np.zeros(k, np.int64)
        offs = np.zeros(k, np.int64)
        lib().orc_node_info(self.h, _p(names), _p(lens), _p(offs))
        return names, lens, offs

    def occ(self, symbol: int, pos: int) -> int:
        return int(lib().orc_occ(self.h, symbol, pos))

    def get_rs(self, pos: int) -> tuple[int, int]:
        rs = int(lib().orc_get_rs(self.h, pos))
        return rs >> 32, rs & 0xFFFFFFFF

    def search(self, pat: bytes) -> tuple[int, int, int]:
        p = _u8(pat)
        sp, ep = C.c_int64(), C.c_int64()
        calls = lib().orc_search(self.h, _p(p), len(p), C.byref(sp), C.byref(ep))
        return sp.value, ep.value, int(calls)

    def search_batch(self, pats: np.ndarray, off: np.ndarray):
        pats = np.ascontiguousarray(pats, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        n = len(off) - 1
        sp = np.zeros(n, np.int64)
        ep = np.zeros(n, np.int64)
        calls = lib().orc_search_batch(self.h, _p(pats), _p(off), n, _p(sp), _p(ep))
        return sp, ep, int(calls)

    def locate(self, row: int) -> int:
        return int(lib().orc_locate(self.h, row))

    def index_find(self, pos: int) -> int:
        return int(lib().orc_index_find(self.h, pos))

    def extract(self, nstr: int, start: int, cap: int) -> np.ndarray:
        """GSSA.extract into a fresh buffer of `cap` bytes: the bytes written (string nstr from `start`)."""
        out = np.zeros(max(cap, 1), np.uint8)
        w = int(lib().orc_extract(self.h, nstr, start, _p(out), cap))
        if w == -(1 << 63):
            raise IndexError(f"String index {nstr} is out of bound")
        if w < 0:
            raise ValueError("newPosition < 0")            # IllegalArgumentException from buf.position(bpos + 1)
        return out[:w].copy()

    def find(self, pat: bytes):
        """GSSA.find: list (per string) of ascending relative positions, or None when no hit."""
        p = _u8(pat)
        sp, ep, _ = self.search(pat)
        if ep < sp:
            return None
        k = ep - sp + 1
        per = np.zeros(max(self.n_strings, 1), np.int64)
        pos = np.zeros(k, np.int64)
        w = lib().orc_find(self.h, _p(p), len(p), _p(per), _p(pos), k)
        res, o = [], 0
        for i in range(self.n_strings):
            c = int(per[i])
            res.append(pos[o:o + c].copy() if c > 0 else None)
            o += c
        assert o == w
        return res


# ---- container level -----------------------------------------------------------------------------
def header_hash(headers) -> int:
    return int(lib().orc_header_hash(_cstrs(headers), len(headers)))


def ref_header(headers, size: int, length: int) -> bytes:
    n = lib().orc_ref_header_len(_cstrs(headers), len(headers))
    out = np.zeros(n, np.uint8)
    w = lib().orc_ref_header_write(_cstrs(headers), len(headers), size, length, _p(out))
    assert w == n
    return out.tobytes()


def ssa_header(headers, length: int) -> bytes:
    out = np.zeros(25, np.uint8)
    lib().orc_ssa_header_write(_cstrs(headers), len(headers), length, _p(out))
    return out.tobytes()


def merge_blocks(lengths, headers):
    """tools/GecoIndex.java:72-98.  Returns list of blocks (file order), each a list of sequence ids."""
    ln = np.ascontiguousarray(lengths, dtype=np.int32)
    n = len(ln)
    bo = np.zeros(n, np.int32)
    pb = np.zeros(n, np.int32)
    nb = lib().orc_merge_blocks(_p(ln), _cstrs(headers), n, _p(bo), _p(pb))
    blocks = [[] for _ in range(nb)]
    for b in range(nb):
        ids = [i for i in range(n) if bo[i] == b]
        ids.sort(key=lambda i: pb[i])
        blocks[b] = ids
    return blocks


def write_files(records, sampling_rate: int = 32):
    """GecoIndex.index + GecozFileWriter: records = [(header, bytes)] -> (gcz bytes, gcx bytes, blocks)."""
    headers = [h for h, _ in records]
    lengths = [len(s) for _, s in records]
    blocks = merge_blocks(lengths, headers)
    gcz, gcx = bytearray(), bytearray()
    for ids in blocks:
        text = b"".join(bytes(records[i][1]) + b"\0" for i in ids)
        hs = [headers[i] for i in ids]
        r = build_block(text, sampling_rate)
        hdr_len = 26 + sum(len(h) + 1 for h in hs)
        gcz += ref_header(hs, hdr_len + len(r["gcz_body"]), len(text)) + r["gcz_body"].tobytes()
        gcx += ssa_header(hs, len(r["gcx_body"])) + r["gcx_body"].tobytes()
    return bytes(gcz), bytes(gcx), blocks
