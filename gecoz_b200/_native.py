"""ctypes binding of libgcz_b200.so — the C ABI declared in include/gcz.h.

There is no CPU fallback: if the shared library is missing this module raises at import of the first
symbol, and every compute entry point fails with GCZ_E_NODEVICE when no CUDA device is visible.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
_SO = _PKG / "libgcz_b200.so"
_CSRC = _PKG / "csrc"

GCZ_OK = 0
GCZ_E_ARG, GCZ_E_NOMEM, GCZ_E_CUDA, GCZ_E_FORMAT, GCZ_E_NODEVICE, GCZ_E_RANGE, GCZ_E_INTERNAL = -1, -2, -3, -4, -5, -6, -7
_ERR_NAMES = {-1: "GCZ_E_ARG", -2: "GCZ_E_NOMEM", -3: "GCZ_E_CUDA", -4: "GCZ_E_FORMAT", -5: "GCZ_E_NODEVICE",
              -6: "GCZ_E_RANGE", -7: "GCZ_E_INTERNAL"}


class GczError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{_ERR_NAMES.get(code, code)}: {message}")
        self.code = code


class GczOutOfMemory(GczError):
    """GCZ_E_NOMEM: the caller may re-queue the block (WriterPoolExecutor.afterExecute semantics)."""


class GczFormatError(GczError):
    """GCZ_E_FORMAT: DataFormatException("invalid index file") in the reference."""


class Shape(C.Structure):
    """struct gcz_shape (include/gcz.h) == HSWTShape (algo/tree/HSWTShape.java:55-87)."""
    _fields_ = [
        ("bit_lengths", C.c_int8 * 256),
        ("codes", C.c_int16 * 256),
        ("n_nodes", C.c_int32),
        ("node_name", C.c_int32 * 256),
        ("node_depth", C.c_int32 * 256),
        ("node_prefix", C.c_int32 * 256),
        ("node_bits", C.c_int64 * 256),
        ("node_offset", C.c_int64 * 256),
        ("table_bytes", C.c_int64),
        ("length", C.c_int64),
        ("size", C.c_int64),
    ]


class BuildTiming(C.Structure):
    _fields_ = [
        ("h2d_ms", C.c_float), ("sort_initial_ms", C.c_float), ("sort_refine_ms", C.c_float),
        ("bwt_hswt_ms", C.c_float), ("ssa_ms", C.c_float), ("d2h_ms", C.c_float), ("total_ms", C.c_float),
        ("refine_rounds", C.c_int32),
        ("radix_launches", C.c_int64), ("radix_elements", C.c_int64),
        ("radix_ms", C.c_float),
        ("kernel_launches", C.c_int64),
        ("symbols_per_key", C.c_int32), ("long_runs", C.c_int32),
        ("unresolved_after_first_sort", C.c_int64),
        ("radix_full_launches", C.c_int64), ("radix_full_ms", C.c_float), ("radix_text_ms", C.c_float),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


EXPORTS = [
    "gcz_init", "gcz_shutdown", "gcz_last_error", "gcz_version", "gcz_device_count", "gcz_set_stream",
    "gcz_shape_from_counts", "gcz_shape_write", "gcz_shape_read", "gcz_ranked_bytes", "gcz_index_size",
    "gcz_count_symbols", "gcz_build_block", "gcz_last_build_timing",
    "gcz_open_block", "gcz_close_block", "gcz_text_length", "gcz_sampling_factor", "gcz_num_strings",
    "gcz_string_ends", "gcz_c_array", "gcz_count_batch", "gcz_locate_rows", "gcz_find_batch", "gcz_extract", "gcz_free",
    "gcz_count_multi", "gcz_count_stats", "gcz_last_query_stats", "gcz_find_multi", "gcz_hits_free",
    "gcz_dbg_set_find_chunk", "gcz_dbg_sort_pairs", "gcz_dbg_suffix_array", "gcz_dbg_ranked_vector", "gcz_dbg_index_wavelet_tree",
]

# include/gcz_file.h — the native host layer (FASTA records, block planning, .gcz/.gcx writer and reader)
FILE_EXPORTS = [
    "gcz_fasta_open", "gcz_fasta_open_buffer", "gcz_fasta_count", "gcz_fasta_record", "gcz_fasta_read", "gcz_fasta_close",
    "gcz_plan_blocks", "gcz_ref_header_length", "gcz_header_hash", "gcz_ref_header_write", "gcz_ssa_header_write",
    "gcz_index_fasta",
    "gcz_reader_open", "gcz_reader_num_blocks", "gcz_reader_block", "gcz_reader_header", "gcz_reader_find",
    "gcz_reader_sampling_factor", "gcz_reader_open_block", "gcz_reader_close",
    "gcz_match", "gcz_gff_search", "gcz_extract_fasta", "gcz_extract_sequence",
]

COUNT_SYMBOLS_FN = C.CFUNCTYPE(C.c_int, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64))
BUILD_BLOCK_FN = C.CFUNCTYPE(C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int32, C.POINTER(Shape), C.c_void_p, C.c_int64,
                             C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p)


class Hits(C.Structure):
    """struct gcz_hits (include/gcz.h)."""
    _fields_ = [("n_hits", C.c_int64), ("block_off", C.POINTER(C.c_int64)), ("pattern", C.POINTER(C.c_int64)),
                ("string", C.POINTER(C.c_int32)), ("position", C.POINTER(C.c_int64))]


class QueryStats(C.Structure):
    """struct gcz_query_stats (include/gcz.h)."""
    _fields_ = [("patterns", C.c_int64), ("blocks", C.c_int64), ("steps", C.c_int64), ("rank_sectors", C.c_int64),
                ("reference_rank_calls", C.c_int64), ("index_bytes", C.c_int64), ("kernel_ms", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Engine(C.Structure):
    """struct gcz_engine (include/gcz_file.h): the per-block device work; null members = the CUDA entry points."""
    _fields_ = [("count_symbols", COUNT_SYMBOLS_FN), ("build_block", BUILD_BLOCK_FN)]


class QueryEngine(C.Structure):
    """struct gcz_query_engine (include/gcz_file.h); null members = the CUDA entry points."""
    _fields_ = [
        ("open_block", C.CFUNCTYPE(C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_void_p))),
        ("close_block", C.CFUNCTYPE(None, C.c_void_p)),
        ("num_strings", C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_int32))),
        ("string_ends", C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_int64))),
        ("find_batch", C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_int64),
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_void_p))),
        ("extract", C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64))),
        ("release", C.CFUNCTYPE(None, C.c_void_p)),
    ]


class IndexReport(C.Structure):
    _fields_ = [("blocks", C.c_int64), ("sequences", C.c_int64), ("symbols", C.c_int64), ("seconds", C.c_double)]


def build(force: bool = False) -> Path:
    """Compile the CUDA sources for sm_100a with nvcc (cross-compiles without a GPU)."""
    srcs = list(_CSRC.glob("*.cu")) + list(_CSRC.glob("*.cuh")) + list(_CSRC.glob("*.cpp")) + list(_CSRC.glob("*.h"))
    srcs.append(_PKG.parent / "include" / "gcz.h")
    stale = force or not _SO.exists() or any(p.stat().st_mtime > _SO.stat().st_mtime for p in srcs)
    if stale:
        subprocess.run(["make", "-C", str(_CSRC), "-j8"], check=True, stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not _SO.exists():
        raise ImportError(f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"(nvcc, sm_100a). gecoz_b200 has no CPU fallback.")
    L = C.CDLL(str(_SO))
    P, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    SP = C.POINTER(Shape)
    sig = {
        "gcz_init": (C.c_int, [C.c_int, P]),
        "gcz_shutdown": (None, []),
        "gcz_last_error": (C.c_char_p, []),
        "gcz_version": (C.c_char_p, []),
        "gcz_device_count": (C.c_int, []),
        "gcz_set_stream": (C.c_int, [C.c_int, P]),
        "gcz_shape_from_counts": (C.c_int, [P, SP]),
        "gcz_shape_write": (i64, [SP, P, i64]),
        "gcz_shape_read": (C.c_int, [P, i64, SP]),
        "gcz_ranked_bytes": (i64, [i64]),
        "gcz_index_size": (i64, [i64, i32]),
        "gcz_count_symbols": (C.c_int, [C.c_int, P, i64, P]),
        "gcz_build_block": (C.c_int, [C.c_int, P, i64, i32, SP, P, i64, P, i64, P, P]),
        "gcz_last_build_timing": (C.c_int, [C.POINTER(BuildTiming)]),
        "gcz_open_block": (C.c_int, [C.c_int, P, i64, i64, P, i64, C.POINTER(P)]),
        "gcz_close_block": (None, [P]),
        "gcz_text_length": (C.c_int, [P, C.POINTER(i64)]),
        "gcz_sampling_factor": (C.c_int, [P, C.POINTER(i32)]),
        "gcz_num_strings": (C.c_int, [P, C.POINTER(i32)]),
        "gcz_string_ends": (C.c_int, [P, P]),
        "gcz_c_array": (C.c_int, [P, P]),
        "gcz_count_batch": (C.c_int, [P, P, P, i64, P, P]),
        "gcz_locate_rows": (C.c_int, [P, P, i64, P]),
        "gcz_find_batch": (C.c_int, [P, P, P, i64, P, C.POINTER(P), C.POINTER(P)]),
        "gcz_extract": (C.c_int, [P, i32, i64, P, i64, C.POINTER(i64)]),
        "gcz_free": (None, [P]),
        "gcz_count_multi": (C.c_int, [P, i32, P, P, i64, P]),
        "gcz_count_stats": (C.c_int, [P, i32, P, P, i64, C.POINTER(QueryStats)]),
        "gcz_last_query_stats": (C.c_int, [C.POINTER(QueryStats)]),
        "gcz_find_multi": (C.c_int, [P, i32, P, P, i64, C.POINTER(Hits)]),
        "gcz_hits_free": (None, [C.POINTER(Hits)]),
        "gcz_dbg_set_find_chunk": (C.c_int, [i64]),
        "gcz_dbg_sort_pairs": (C.c_int, [C.c_int, P, P, i64, i32, i32]),
        "gcz_dbg_suffix_array": (C.c_int, [C.c_int, P, i64, P]),
        "gcz_dbg_ranked_vector": (C.c_int, [C.c_int, P, i64, P]),
        "gcz_dbg_index_wavelet_tree": (C.c_int, [C.c_int, P, i64, P]),
    }
    assert sorted(sig) == sorted(EXPORTS)
    PP = C.POINTER(C.c_char_p)
    sig.update({
        "gcz_fasta_open": (C.c_int, [C.c_char_p, C.POINTER(P)]),
        "gcz_fasta_open_buffer": (C.c_int, [P, i64, C.POINTER(P)]),
        "gcz_fasta_count": (i64, [P]),
        "gcz_fasta_record": (C.c_int, [P, i64, PP, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]),
        "gcz_fasta_read": (C.c_int, [P, i64, P, i64]),
        "gcz_fasta_close": (None, [P]),
        "gcz_plan_blocks": (i64, [P, PP, i64, P, P]),
        "gcz_ref_header_length": (i64, [PP, i32]),
        "gcz_header_hash": (i64, [PP, i32]),
        "gcz_ref_header_write": (i64, [PP, i32, i64, i64, P, i64]),
        "gcz_ssa_header_write": (i64, [PP, i32, i64, P]),
        "gcz_index_fasta": (C.c_int, [P, C.c_char_p, C.c_char_p, i32, i32, P, C.POINTER(Engine), C.POINTER(IndexReport)]),
        "gcz_reader_open": (C.c_int, [C.c_char_p, C.POINTER(P)]),
        "gcz_reader_num_blocks": (i32, [P]),
        "gcz_reader_block": (C.c_int, [P, i32, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]),
        "gcz_reader_header": (C.c_char_p, [P, i32, i32]),
        "gcz_reader_find": (C.c_int, [P, C.c_char_p, C.POINTER(i32), C.POINTER(i32)]),
        "gcz_reader_sampling_factor": (i32, [P]),
        "gcz_reader_open_block": (C.c_int, [P, i32, C.c_int, C.POINTER(P)]),
        "gcz_reader_close": (None, [P]),
        "gcz_match": (C.c_int, [P, C.c_int, C.c_char_p, P, i64, i32, C.POINTER(QueryEngine), C.POINTER(P), C.POINTER(i64)]),
        "gcz_gff_search": (C.c_int, [P, C.c_int, P, i64, C.POINTER(QueryEngine), C.POINTER(P), C.POINTER(i64)]),
        "gcz_extract_fasta": (C.c_int, [P, C.c_int, C.c_char_p, C.POINTER(QueryEngine), C.POINTER(i64)]),
        "gcz_extract_sequence": (C.c_int, [P, C.c_int, C.c_char_p, i64, i64, C.c_char_p, C.POINTER(QueryEngine), C.POINTER(i64)]),
    })
    assert sorted(sig) == sorted(EXPORTS + FILE_EXPORTS)
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def check(rc: int) -> int:
    if rc is not None and rc < 0:
        msg = lib().gcz_last_error().decode("utf-8", "replace")
        if rc == GCZ_E_NOMEM:
            raise GczOutOfMemory(rc, msg)
        if rc == GCZ_E_FORMAT:
            raise GczFormatError(rc, msg)
        raise GczError(rc, msg)
    return rc


def ptr(a) -> C.c_void_p:
    """Address of a numpy array, a torch tensor (host or CUDA), a ctypes buffer, or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.cast(a, C.c_void_p)
