// GPU suffix array construction (see suffix_sort.cu).
#pragma once

#include "radix_sort.cuh"

namespace gcz {

struct SuffixSortStats {
    float   initial_ms = 0, refine_ms = 0;
    int     rounds = 0;
    int     symbols_per_key = 0;
    int64_t radix_passes = 0, radix_elements = 0;
    float   radix_ms = 0;
    int64_t radix_full_passes = 0;     // digit passes over all n pairs with array input: the roofline kernel
    float   radix_full_ms = 0;
    float   radix_text_ms = 0;         // the first digit pass (reads the text)
    int64_t unresolved_after_first_sort = 0;
    int     long_runs = 0;
};

size_t suffix_sort_workspace_bytes(int64_t n);

// d_text: n bytes on the device; counts: byte histogram of the text (host); d_sa: n x u32 on the device (output).
// Scratch comes from `arena` (released before returning).
// carry_shift (optional): in, nonzero = the caller accepts SA entries that carry the dense code (1..sigma) of the
// preceding text symbol above the position; out, the bit the code starts at, or 0 when the entries are plain.
int suffix_sort(DeviceCtx* ctx, cudaStream_t st, const uint8_t* d_text, int64_t n, const int64_t counts[256],
                uint32_t* d_sa, Arena& arena, SuffixSortStats* stats, int* carry_shift = nullptr);

}  // namespace gcz
