// Query-side structures (see query.cu).
#pragma once

#include "radix_sort.cuh"

#include <memory>
#include <vector>

namespace gcz {

// Everything a query kernel needs besides the rank sectors; copied to shared memory by every CTA.
struct QueryTables {
    int64_t  c[256];                // GSSA.c: symbols smaller than i
    int64_t  n;                     // text length
    uint64_t node_sector0[256];     // node (file order) -> first rank sector
    uint64_t marker_sector0;
    uint64_t iwt_sector0[64];       // IndexWaveletTree level (bit) -> first rank sector
    int64_t  iwt_m;
    int32_t  iwt_levels;
    int32_t  sampling_factor;
    int32_t  n_nodes;
    int32_t  pad_;
    int16_t  child[256][2];         // node, bit -> node (>= 0) or ~symbol (< 0)
    uint16_t code[256];             // byte -> code bits
    uint8_t  len[256];              // byte -> code length (0 = absent)
    uint8_t  node_of[256][16];      // byte, depth -> node on the symbol's path
    // (sp, ep) of every string of kmer_k symbols out of A, C, G, T, as the backward search leaves them (x > y: the search of
    // that string fails on the way): built at open by that very search, looked up for the last kmer_k symbols of a pattern
    const uint2* kmer;              // 4^kmer_k entries, index = the symbols' 2-bit codes, first symbol highest; null: no table
    int32_t  kmer_k;
    int32_t  pad2_;
    uint8_t  code2[256];            // byte -> 0..3 for A, C, G, T, 0xFF otherwise
};

}  // namespace gcz

struct gcz_index {
    gcz::DeviceCtx*   ctx = nullptr;
    int64_t           n = 0;
    int32_t           sampling_factor = 0;
    gcz_shape         shape;
    uint32_t*         d_sectors = nullptr;
    size_t            sector_bytes = 0;
    gcz::QueryTables* d_tables = nullptr;
    uint2*            d_kmer = nullptr;
    int64_t           c[256];
    std::vector<int64_t> e;          // sorted separator positions
};

namespace gcz {

int  open_block(DeviceCtx* ctx, const uint8_t* gcz_body, int64_t body_len, int64_t text_len,
                const uint8_t* gcx_body, int64_t gcx_len, gcz_index** out);
void close_block(gcz_index* idx);
int  count_batch(gcz_index* idx, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats, int64_t* sp, int64_t* ep);
int  locate_rows(gcz_index* idx, const int64_t* rows, int64_t n_rows, int64_t* positions);
int  extract(gcz_index* idx, int32_t nstr, int64_t from, uint8_t* out, int64_t cap, int64_t* written);
int  find_batch(gcz_index* idx, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats,
                int64_t* per_string_counts, int64_t** positions, int64_t** pos_off);
int  count_multi(gcz_index* const* blocks, int32_t n_blocks, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats, int64_t* totals);
int  count_stats(gcz_index* const* blocks, int32_t n_blocks, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats, gcz_query_stats* out);
int  last_query_stats(gcz_query_stats* out);
void set_find_chunk(int64_t occurrences);
int  find_multi(gcz_index* const* blocks, int32_t n_blocks, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats, gcz_hits* out);

}  // namespace gcz
