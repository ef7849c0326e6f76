"""gecoz_b200 — B200-native FM-index engine behind gecoz's nova-algo API.

Layout: csrc/ holds the CUDA kernels and the C ABI (include/gcz.h -> libgcz_b200.so); the Python modules are
the host-side mirror of the reference's Java interface for this path (GecozFileWriter / GecozFileReader /
GSSA / GecoIndex), because no JVM exists in this environment.  Nothing here falls back to the CPU.
"""
from ._native import GczError, GczFormatError, GczOutOfMemory, Shape, build, lib
from .gecoz_file import (GecozFileReader, GecozFileWriter, GecozRefBlockHeader, GecozSSABlockHeader, build_block,
                         index_size, shape_from_counts, symbol_counts)
from .gssa import GSSA, count_stats, count_totals, find_multi, find_total, last_query_stats, pack_patterns
from .geco_index import FastaSequence, GecozRefBlock, index, index_records, merge_blocks, read_fasta
from . import geco_match, geco_read, native_file, sharding

__all__ = [
    "GczError", "GczFormatError", "GczOutOfMemory", "Shape", "build", "lib",
    "GecozFileReader", "GecozFileWriter", "GecozRefBlockHeader", "GecozSSABlockHeader", "build_block", "index_size",
    "shape_from_counts", "symbol_counts", "GSSA", "pack_patterns", "count_totals", "count_stats", "last_query_stats", "find_multi", "find_total",
    "FastaSequence", "GecozRefBlock", "index", "index_records", "merge_blocks", "read_fasta", "sharding", "geco_match", "geco_read", "native_file",
]
