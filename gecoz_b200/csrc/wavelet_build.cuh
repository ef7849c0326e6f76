// BWT / HSWT / sampled-SA index builders (see wavelet_build.cu).
#pragma once

#include "device.cuh"

namespace gcz {

struct WaveletStats {
    float bwt_hswt_ms = 0, ssa_ms = 0;
};

size_t wavelet_workspace_bytes(int64_t n, int sampling_factor);

// Inputs on the device: text (n bytes), suffix array (n x u32).  Outputs on the device: d_bwt (n bytes),
// d_gcz_body (shape->size bytes: shape table + ranked HSWT nodes), d_gcx_body (index_size bytes: ranked marker
// vector + IndexWaveletTree levels).  `shape` must come from shape_from_counts on this text's histogram.
// carry_shift: as returned by suffix_sort (SA entries carry their BWT symbol); clean_sa: leave plain positions in d_sa.
int build_wavelet_structures(DeviceCtx* ctx, cudaStream_t st, const uint8_t* d_text, uint32_t* d_sa, int carry_shift, bool clean_sa,
                             int64_t n, const gcz_shape* shape, int sampling_factor, uint8_t* d_bwt,
                             uint8_t* d_gcz_body, uint8_t* d_gcx_body, Arena& arena, WaveletStats* stats,
                             uint8_t* h_gcz_out = nullptr, cudaStream_t copy_stream = nullptr, cudaEvent_t gcz_copied = nullptr);
// h_gcz_out (optional, host): the finished .gcz body is copied there on copy_stream while the index is still being
// built; gcz_copied is recorded on copy_stream after that copy (the caller waits for it).

// stage hooks (parity tests)
int ranked_vector_from_bits(DeviceCtx* ctx, cudaStream_t st, const uint8_t* d_bits, int64_t len, uint8_t* d_out, Arena& arena);
int index_wavelet_tree_from_values(DeviceCtx* ctx, cudaStream_t st, const uint32_t* d_vals, int64_t m, uint8_t* d_out, Arena& arena);

}  // namespace gcz
